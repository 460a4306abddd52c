// csg_raycaster.hpp — C++ mirror of the reference's host interface for the raycast path, over the C ABI.
//
// A maintainer of the reference swaps
//     #include "Graphics/RayCasting/Raycaster.cuh"      ->  #include "csg_raycaster.hpp"   (and links libcsg_b200.so)
// and keeps calling the same names with the same argument meaning and error behaviour:
//     CSGTree::Parse(text)                    (RayCasting/CSGTree/CSGTree.cuh:53)   throws std::invalid_argument
//     Raycaster::ChangeSize(w, h, tree)       (RayCasting/Raycaster.cuh:22)
//     Raycaster::Raycast(devPBO, cam, light)  (RayCasting/Raycaster.cuh:23)         float4 device buffer, synchronous
//     Raycaster::CleanUp()                    (RayCasting/Raycaster.cuh:24)
//     Camera / DirectionalLight               (RenderManager/Camera/Camera.h, RenderManager/DirectionalLight.h)
// Header-only; no CUDA headers needed by the caller.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

#include <vector>

#include "../../include/csg_b200.h"

namespace csg_b200 {

struct float3_t { float x, y, z; };

class Camera {  // same public surface as the reference's Camera (Camera.h:6-53)
  public:
    float x, y, z;
    float rotX, rotY;
    float fov;
    float forward[3], right[3], up[3];

    Camera() { csg_camera c; csg_camera_default(&c); load(c); }
    void setPosition(float px, float py, float pz) { x = px; y = py; z = pz; }
    void setRotation(float pitch, float yaw) { csg_camera c = store(); csg_camera_set(&c, x, y, z, pitch, yaw); c.fov = fov; load(c); }
    void move(float f, float r, float u)
    {  // Camera.h:36-41
        x += forward[0] * f + right[0] * r;
        y += forward[1] * f + right[1] * r + u;
        z += forward[2] * f + right[2] * r;
    }
    void rotate(float dPitch, float dYaw) { setRotation(rotX + dPitch, rotY + dYaw); }
    void setFOV(float degrees) { csg_camera c = store(); csg_camera_set_fov_degrees(&c, degrees); fov = c.fov; }
    csg_camera store() const
    {
        csg_camera c;
        static_assert(sizeof(csg_camera) == 60, "Camera layout");
        std::memcpy(&c, this, sizeof c);   // identical field order
        return c;
    }

  private:
    void load(const csg_camera& c) { std::memcpy(this, &c, sizeof c); }
};
static_assert(sizeof(Camera) == 60, "Camera must stay layout-compatible with csg_camera");

struct DirectionalLight {  // DirectionalLight.h:8-18
    float polar, azimuth;
    DirectionalLight() { csg_light l; csg_light_default(&l); polar = l.polar; azimuth = l.azimuth; }
    float3_t getLightDir() const
    {
        csg_light l{polar, azimuth};
        float d[3];
        csg_light_direction(&l, d);
        return float3_t{d[0], d[1], d[2]};
    }
};

class CSGTree {  // value type like the reference's (two std::vectors there, one handle here)
  public:
    CSGTree() = default;
    CSGTree(const CSGTree&) = delete;
    CSGTree& operator=(const CSGTree&) = delete;
    CSGTree(CSGTree&& o) noexcept : scene_(o.scene_) { o.scene_ = nullptr; }
    CSGTree& operator=(CSGTree&& o) noexcept { if (this != &o) { csg_free_scene(scene_); scene_ = o.scene_; o.scene_ = nullptr; } return *this; }
    ~CSGTree() { csg_free_scene(scene_); }

    // throws std::invalid_argument with the reference's message (CSGTree.cu:24, 68, 104-108, 140, 147)
    static CSGTree Parse(const std::string& text)
    {
        CSGTree t;
        if (csg_parse_scene(text.data(), text.size(), &t.scene_) != CSG_OK) throw std::invalid_argument(csg_last_error());
        return t;
    }
    const csg_scene* handle() const { return scene_; }
    csg_scene* handle() { return scene_; }

  private:
    csg_scene* scene_ = nullptr;
};

class Raycaster {
  public:
    Raycaster() = default;
    Raycaster(const Raycaster&) = delete;
    Raycaster& operator=(const Raycaster&) = delete;
    ~Raycaster() { CleanUp(); }

    // Raycaster::ChangeSize (Raycaster.cu:3-21).  The reference exits the process on a CUDA error; this throws.
    void ChangeSize(int newWidth, int newHeight, const CSGTree& tree, int gpus = 1)
    {
        CleanUp();
        if (csg_upload(tree.handle(), newWidth, newHeight, gpus, &ctx_) != CSG_OK) throw std::runtime_error(csg_last_error());
    }
    // Raycaster::Raycast (Raycaster.cu:23-34): devPBO = width*height float4 in device memory (the mapped GL PBO).
    void Raycast(void* devPBO_float4, const Camera& cam, const DirectionalLight& light)
    {
        csg_camera c = cam.store();
        csg_light l{light.polar, light.azimuth};
        if (csg_render_f32(ctx_, &c, &l, static_cast<float*>(devPBO_float4)) != CSG_OK) throw std::runtime_error(csg_last_error());
    }
    // RGBA8 form (host or device pointer) — what a headless caller wants.
    void RaycastRGBA8(uint8_t* rgba8, const Camera& cam, const DirectionalLight& light)
    {
        csg_camera c = cam.store();
        csg_light l{light.polar, light.azimuth};
        if (csg_render(ctx_, &c, &l, rgba8) != CSG_OK) throw std::runtime_error(csg_last_error());
    }
    void CleanUp()
    {  // idempotent, like the reference's `alloced` guard (Raycaster.cu:36-45)
        csg_free_context(ctx_);
        ctx_ = nullptr;
    }
    // A camera path in one call (the application's frame loop over a moving camera, Application.cpp:39-57): n frames of
    // RGBA8, frame k at rgba8 + k*width*height*4 (host or device pointer), pipelined over two frame slots.
    void RaycastBatch(const Camera* cams, int n, const DirectionalLight& light, uint8_t* rgba8)
    {
        std::vector<csg_camera> cs(static_cast<size_t>(n));
        for (int k = 0; k < n; ++k) cs[static_cast<size_t>(k)] = cams[k].store();
        csg_light l{light.polar, light.azimuth};
        if (csg_render_batch(ctx_, cs.data(), n, &l, rgba8) != CSG_OK) throw std::runtime_error(csg_last_error());
    }
    // Not in the reference (its kernel has one ray per pixel, no culling options): samples per axis (1, 2 or 4 rays per pixel
    // side: csg_set_supersampling), the per-tile tree pruning mode (0 off, 1 default, 2 tree walk: csg_set_pruning) and the view
    // cache for a static camera with a moving light (csg_set_view_cache).  Call after ChangeSize.
    void SetSupersampling(int samplesPerAxis) { if (csg_set_supersampling(ctx_, samplesPerAxis) != CSG_OK) throw std::runtime_error(csg_last_error()); }
    void SetPruning(int mode) { if (csg_set_pruning(ctx_, mode) != CSG_OK) throw std::runtime_error(csg_last_error()); }
    void SetViewCache(bool on) { if (csg_set_view_cache(ctx_, on ? 1 : 0) != CSG_OK) throw std::runtime_error(csg_last_error()); }
    // the CUDA stream (cudaStream_t) frames are enqueued on — the place of the default stream Raycast shares with the GL map / unmap
    // (RenderManager.cpp:58-84): work the caller queues there is ordered against the frames on the device
    void* Stream() { void* s = nullptr; if (csg_stream(ctx_, &s) != CSG_OK) throw std::runtime_error(csg_last_error()); return s; }
    csg_context* context() { return ctx_; }

  private:
    csg_context* ctx_ = nullptr;
};

}  // namespace csg_b200
