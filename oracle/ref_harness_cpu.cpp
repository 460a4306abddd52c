// TEST INFRASTRUCTURE — not part of the product.
//
// Host build of the UNMODIFIED reference hot path.  This file contains no
// reference code: it #includes the reference's own RaycastingKernels.cu
// (path given by -DREF_KERNELS_CU; the Makefile points it at a build-time temp
// copy of /root/reference/.../RaycastingKernels.cu whose only change is the
// missing `return true;` at the end of cylinderHit, RaycastingKernels.cu:335-336,
// without which g++ -O1+ treats the fall-through as unreachable) and calls
// RaycastKernel / LightningKernel once per pixel as ordinary functions.
//
// How that works without nvcc: outside __CUDACC__ the CUDA headers define
// __global__/__device__ away, and device_launch_parameters.h declares
// threadIdx/blockIdx/blockDim with the storage class __STORAGE__, which the
// Makefile pre-defines as `extern thread_local`, so each OpenMP thread can set
// "its" thread index before calling the kernel body.
#include <cuda_runtime.h>
#include <device_launch_parameters.h>
#include <omp.h>
#include <chrono>
#include <cstring>
#include <string>

extern "C" {
thread_local uint3 threadIdx;
thread_local uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;
thread_local int warpSize = 32;
}

#include REF_KERNELS_CU  // the reference's RaycastingKernels.cu (+ its headers)

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
    Camera cam;  // reference defaults: pos (0,0,5), fov 90*3.14159/180
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

extern "C" {

int refcpu_tree_info(const char* text, int* n_nodes, int* n_prims, char* err, int errlen)
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

// nodes44: n_nodes * 44 bytes (CSGNode), prims48: n_prims * 48 bytes (Primitive)
int refcpu_tree_dump(const char* text, void* nodes44, void* prims48)
{
    static_assert(sizeof(CSGNode) == 44 && sizeof(Primitive) == 48, "layout");
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

// out15 = x,y,z, rotX,rotY, fov, forward[3], right[3], up[3]  (Camera.h:9-16)
void refcpu_camera(const ref_view* v, float* out15)
{
    Camera cam = make_camera(v);
    static_assert(sizeof(Camera) == 60, "layout");
    std::memcpy(out15, &cam, 60);
}

void refcpu_light_dir(const ref_view* v, float* out3)
{
    DirectionalLight l = make_light(v);
    float3 d = l.getLightDir();
    out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}

int refcpu_max_threads(void) { return omp_get_max_threads(); }

// Renders rows y0, y0+row_step, ... < y1 of the frame (row 0 = bottom scanline, as the reference).
// Output arrays are full-frame (w*h), only the rendered rows are written; any may be NULL.
// rgba = 4 floats per pixel exactly as LightningKernel writes them.
int refcpu_render(const char* text, const ref_view* v, int y0, int y1, int row_step, int nthreads,
                  uint8_t* hit, int32_t* prim, float* t, float* rgba,
                  double* seconds, char* err, int errlen)
{
    CSGTree tree;
    try {
        tree = CSGTree::Parse(text);
    } catch (const std::exception& e) {
        fill_err(err, errlen, e.what());
        return 1;
    }
    CudaCSGTree ct;
    ct.nodes = tree.nodes.data();
    ct.primitives = tree.primitives.primitives.data();
    Camera cam = make_camera(v);
    DirectionalLight light = make_light(v);
    float3 lightDir = light.getLightDir();
    const int w = v->width, h = v->height;
    if (y0 < 0) y0 = 0;
    if (y1 > h) y1 = h;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    if (row_step < 1) row_step = 1;
    const int nrows = (y1 - y0 + row_step - 1) / row_step;

    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
    for (int yi = 0; yi < nrows; ++yi) {
        const int y = y0 + yi * row_step;
        blockDim = dim3(1, 1, 1);
        blockIdx = uint3{0, 0, 0};
        for (int x = 0; x < w; ++x) {
            RayHit rh;
            float4 px;
            // one "thread" of the reference grid; hits/output pointers are offset so that
            // the kernel's own pixelIdx = y*(int)width + x lands on element 0 of our locals
            threadIdx = uint3{(unsigned)x, (unsigned)y, 0};
            long idx = (long)y * w + x;
            RaycastKernel(cam, ct, &rh - idx, (float)w, (float)h);
            LightningKernel(cam, &rh - idx, ct.primitives, &px - idx, lightDir, (float)w, (float)h);
            if (hit) hit[idx] = rh.hit ? 1 : 0;
            if (prim) prim[idx] = rh.hit ? rh.primitiveIdx : -1;
            if (t) t[idx] = rh.hit ? rh.t : -1.0f;
            if (rgba) { rgba[4 * idx] = px.x; rgba[4 * idx + 1] = px.y; rgba[4 * idx + 2] = px.z; rgba[4 * idx + 3] = px.w; }
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}

}  // extern "C"
