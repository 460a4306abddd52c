// TEST INFRASTRUCTURE — not part of the product.
//
// GPU build of the UNMODIFIED reference hot path: this file contains no reference
// code.  oracle/Makefile compiles it with nvcc -arch=sm_100 together with the
// reference's own RaycastingKernels.cu, Raycaster.cu, CSGTree.cu and Camera.cpp,
// in place from /root/reference, into oracle/_ref/libref_gpu.so.
//
// It launches RaycastKernel + LightningKernel with the reference's grid/block
// (Raycaster.cuh:7-8, Raycaster.cu:10-11), copies RayHit[] and the float4 image
// back, and times (a) the two kernels back to back with CUDA events and (b) the
// shipped Raycaster::Raycast (two launches + two cudaDeviceSynchronize,
// Raycaster.cu:23-34).  This is the primary golden of BASELINE.json's north_star.
#include <cuda_runtime.h>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "RayCasting/Raycaster.cuh"   // -I <reference>/CSGRayCasting/Graphics

#include "ref_harness.h"

static void fill_err(char* err, int errlen, const char* msg)
{
    if (err && errlen > 0) {
        std::strncpy(err, msg, errlen - 1);
        err[errlen - 1] = 0;
    }
}

static Camera make_camera(const ref_view* v)
{
    Camera cam;
    cam.setPosition(v->pos[0], v->pos[1], v->pos[2]);
    cam.setRotation(v->pitch, v->yaw);
    if (v->fov > 0) cam.fov = v->fov;
    return cam;
}

static DirectionalLight make_light(const ref_view* v)
{
    DirectionalLight l;
    if (v->polar < 1e9f) {
        l.polar = v->polar;
        l.azimuth = v->azimuth;
    }
    return l;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fill_err(err, errlen, cudaGetErrorString(e_)); return 2; } } while (0)

extern "C" {

int refgpu_tree_info(const char* text, int* n_nodes, int* n_prims, char* err, int errlen)
{
    try {
        CSGTree tree = CSGTree::Parse(text);
        *n_nodes = (int)tree.nodes.size();
        *n_prims = (int)tree.primitives.primitives.size();
        return 0;
    } catch (const std::exception& e) {
        fill_err(err, errlen, e.what());
        return 1;
    }
}

int refgpu_tree_dump(const char* text, void* nodes44, void* prims48)
{
    try {
        CSGTree tree = CSGTree::Parse(text);
        std::memcpy(nodes44, tree.nodes.data(), tree.nodes.size() * sizeof(CSGNode));
        std::memcpy(prims48, tree.primitives.primitives.data(),
                    tree.primitives.primitives.size() * sizeof(Primitive));
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}

void refgpu_camera(const ref_view* v, float* out15)
{
    Camera cam = make_camera(v);
    std::memcpy(out15, &cam, 60);
}

void refgpu_light_dir(const ref_view* v, float* out3)
{
    DirectionalLight l = make_light(v);
    float3 d = l.getLightDir();
    out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}

// Renders one frame with the reference kernels.  Outputs (host, any may be NULL):
//   hit[w*h] (RayHit.hit), prim[w*h] (RayHit.primitiveIdx, -1 on miss), t[w*h] (RayHit.t, -1 on miss),
//   rgba[w*h*4] float, exactly what LightningKernel wrote.
// Timing (any may be NULL): iters >= 1 timed repetitions after `warmup` untimed ones;
//   ms_kernels[iters]  = CUDA-event time of RaycastKernel+LightningKernel launched back to back,
//   ms_shipped[iters]  = CUDA-event time around Raycaster::Raycast as shipped.
/* Diagnostic: RayHit.position and RayHit.normal (Ray.cuh:24-34) of every pixel as the reference's RaycastKernel stores them
 * (6 floats per pixel; untouched where the ray misses). */
int refgpu_render_details(const char* text, const ref_view* v, float* pos_normal, char* err, int errlen)
{
    CSGTree tree;
    try {
        tree = CSGTree::Parse(text);
    } catch (const std::exception& e) {
        fill_err(err, errlen, e.what());
        return 1;
    }
    const int w = v->width, h = v->height;
    const size_t npx = (size_t)w * h;
    Camera cam = make_camera(v);
    CudaCSGTree ct;
    RayHit* dHits = nullptr;
    CK(cudaMalloc(&ct.nodes, tree.nodes.size() * sizeof(CSGNode)));
    CK(cudaMalloc(&ct.primitives, tree.primitives.primitives.size() * sizeof(Primitive)));
    CK(cudaMemcpy(ct.nodes, tree.nodes.data(), tree.nodes.size() * sizeof(CSGNode), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ct.primitives, tree.primitives.primitives.data(),
                  tree.primitives.primitives.size() * sizeof(Primitive), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dHits, npx * sizeof(RayHit)));
    CK(cudaMemset(dHits, 0, npx * sizeof(RayHit)));
    dim3 block(BLOCKXSIZE, BLOCKYSIZE);
    dim3 grid((w + block.x - 1) / block.x, (h + block.y - 1) / block.y);
    RaycastKernel<<<grid, block>>>(cam, ct, dHits, (float)w, (float)h);
    CK(cudaDeviceSynchronize());
    std::vector<RayHit> hh(npx);
    CK(cudaMemcpy((void*)hh.data(), dHits, npx * sizeof(RayHit), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < npx; ++i) {
        if (!hh[i].hit) continue;
        float* o = pos_normal + 6 * i;
        o[0] = hh[i].position.x; o[1] = hh[i].position.y; o[2] = hh[i].position.z;
        o[3] = hh[i].normal.x; o[4] = hh[i].normal.y; o[5] = hh[i].normal.z;
    }
    cudaFree(ct.nodes);
    cudaFree(ct.primitives);
    cudaFree(dHits);
    return 0;
}

int refgpu_render(const char* text, const ref_view* v,
                  uint8_t* hit, int32_t* prim, float* t, float* rgba,
                  int warmup, int iters, float* ms_kernels, float* ms_shipped,
                  char* err, int errlen)
{
    CSGTree tree;
    try {
        tree = CSGTree::Parse(text);
    } catch (const std::exception& e) {
        fill_err(err, errlen, e.what());
        return 1;
    }
    const int w = v->width, h = v->height;
    const size_t npx = (size_t)w * h;
    Camera cam = make_camera(v);
    DirectionalLight light = make_light(v);

    CudaCSGTree ct;
    RayHit* dHits = nullptr;
    float4* dOut = nullptr;
    CK(cudaMalloc(&ct.nodes, tree.nodes.size() * sizeof(CSGNode)));
    CK(cudaMalloc(&ct.primitives, tree.primitives.primitives.size() * sizeof(Primitive)));
    CK(cudaMemcpy(ct.nodes, tree.nodes.data(), tree.nodes.size() * sizeof(CSGNode), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ct.primitives, tree.primitives.primitives.data(),
                  tree.primitives.primitives.size() * sizeof(Primitive), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dHits, npx * sizeof(RayHit)));
    CK(cudaMalloc(&dOut, npx * sizeof(float4)));
    CK(cudaMemset(dHits, 0, npx * sizeof(RayHit)));

    dim3 block(BLOCKXSIZE, BLOCKYSIZE);
    dim3 grid((w + block.x - 1) / block.x, (h + block.y - 1) / block.y);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    if (iters < 1) iters = 1;
    for (int i = 0; i < warmup + iters; ++i) {
        CK(cudaEventRecord(e0));
        RaycastKernel<<<grid, block>>>(cam, ct, dHits, (float)w, (float)h);
        LightningKernel<<<grid, block>>>(cam, dHits, ct.primitives, dOut, light.getLightDir(), (float)w, (float)h);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= warmup && ms_kernels) ms_kernels[i - warmup] = ms;
    }

    if (hit || prim || t) {
        std::vector<RayHit> hh(npx);
        CK(cudaMemcpy((void*)hh.data(), dHits, npx * sizeof(RayHit), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < npx; ++i) {
            if (hit) hit[i] = hh[i].hit ? 1 : 0;
            if (prim) prim[i] = hh[i].hit ? hh[i].primitiveIdx : -1;
            if (t) t[i] = hh[i].hit ? hh[i].t : -1.0f;
        }
    }
    if (rgba) CK(cudaMemcpy(rgba, dOut, npx * sizeof(float4), cudaMemcpyDeviceToHost));

    if (ms_shipped) {
        Raycaster rc;
        rc.ChangeSize(w, h, tree);
        for (int i = 0; i < warmup + iters; ++i) {
            CK(cudaEventRecord(e0));
            rc.Raycast(dOut, cam, light);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= warmup) ms_shipped[i - warmup] = ms;
        }
        rc.CleanUp();
    }

    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dHits);
    cudaFree(dOut);
    cudaFree(ct.nodes);
    cudaFree(ct.primitives);
    return 0;
}

/* The unmodified reference kernels at the FULL size of `v` (e.g. the 30720 x 17280 virtual grid of BASELINE.json configs[4]:
 * 19 GB of RayHit + 8.5 GB of float4 on the device), of which only the window [x0, x0 + ww) x [y0, y0 + wh) is copied out:
 * rgba (ww * wh float4), hit / prim / t (ww * wh each; any may be NULL).  Lets a test compare a supersampled frame with the
 * reference sample by sample without moving the whole grid to the host. */
int refgpu_render_window(const char* text, const ref_view* v, int x0, int y0, int ww, int wh,
                         uint8_t* hit, int32_t* prim, float* t, float* rgba, float* ms_kernels, char* err, int errlen)
{
    CSGTree tree;
    try {
        tree = CSGTree::Parse(text);
    } catch (const std::exception& e) {
        fill_err(err, errlen, e.what());
        return 1;
    }
    const int w = v->width, h = v->height;
    if (x0 < 0 || y0 < 0 || ww < 1 || wh < 1 || x0 + ww > w || y0 + wh > h) { fill_err(err, errlen, "window outside the frame"); return 2; }
    const size_t npx = (size_t)w * h;
    Camera cam = make_camera(v);
    DirectionalLight light = make_light(v);
    CudaCSGTree ct;
    RayHit* dHits = nullptr;
    float4* dOut = nullptr;
    CK(cudaMalloc(&ct.nodes, tree.nodes.size() * sizeof(CSGNode)));
    CK(cudaMalloc(&ct.primitives, tree.primitives.primitives.size() * sizeof(Primitive)));
    CK(cudaMemcpy(ct.nodes, tree.nodes.data(), tree.nodes.size() * sizeof(CSGNode), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ct.primitives, tree.primitives.primitives.data(),
                  tree.primitives.primitives.size() * sizeof(Primitive), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dHits, npx * sizeof(RayHit)));
    CK(cudaMalloc(&dOut, npx * sizeof(float4)));
    CK(cudaMemset(dHits, 0, npx * sizeof(RayHit)));
    dim3 block(BLOCKXSIZE, BLOCKYSIZE);
    dim3 grid((w + block.x - 1) / block.x, (h + block.y - 1) / block.y);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    RaycastKernel<<<grid, block>>>(cam, ct, dHits, (float)w, (float)h);
    LightningKernel<<<grid, block>>>(cam, dHits, ct.primitives, dOut, light.getLightDir(), (float)w, (float)h);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    if (ms_kernels) CK(cudaEventElapsedTime(ms_kernels, e0, e1));
    if (rgba)
        CK(cudaMemcpy2D(rgba, (size_t)ww * sizeof(float4), dOut + (size_t)y0 * w + x0, (size_t)w * sizeof(float4), (size_t)ww * sizeof(float4), wh,
                        cudaMemcpyDeviceToHost));
    if (hit || prim || t) {
        std::vector<RayHit> hh((size_t)ww * wh);
        CK(cudaMemcpy2D((void*)hh.data(), (size_t)ww * sizeof(RayHit), dHits + (size_t)y0 * w + x0, (size_t)w * sizeof(RayHit), (size_t)ww * sizeof(RayHit), wh,
                        cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < hh.size(); ++i) {
            if (hit) hit[i] = hh[i].hit ? 1 : 0;
            if (prim) prim[i] = hh[i].hit ? hh[i].primitiveIdx : -1;
            if (t) t[i] = hh[i].hit ? hh[i].t : -1.0f;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dHits);
    cudaFree(dOut);
    cudaFree(ct.nodes);
    cudaFree(ct.primitives);
    return 0;
}

}  // extern "C"
