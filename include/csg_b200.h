/* csg_b200 — C ABI of the B200-native CSG ray caster (libcsg_b200.so).
 *
 * Drop-in boundary for the per-pixel raycast + Phong path of
 * Zumi002/CUDA-CSG-Tree-Raycasting.  The reference has no FFI; its seam is the C++ class
 * `Raycaster` plus `CSGTree::Parse`.  Each entry point below names the
 * reference interface it replaces (paths relative to CSGRayCasting/Graphics/).
 *
 * Conventions: plain C, opaque handles, int status codes (0 = CSG_OK), no exceptions and no
 * exit() across the boundary (the reference's gpuErrchk exits the process,
 * RayCasting/Kernels/RaycastingKernels.cuh:22-31).  A handle may be used by one thread at a
 * time.  csg_last_error() is thread-local.  There is no CPU fallback: every render entry
 * point fails with CSG_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Image convention (same as the reference): pixel index = y*width + x, row 0 is the BOTTOM
 * scanline; RGBA8 byte = (int)(clamp(c,0,1)*255 + 0.5), alpha = 255 (Q12).
 */
#ifndef CSG_B200_H
#define CSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct csg_scene csg_scene;     /* host: parsed + flattened CSG tree */
typedef struct csg_context csg_context; /* device(s): uploaded tree, framebuffers, streams */

enum csg_status {
    CSG_OK = 0,
    CSG_ERR_PARSE = 1,     /* scene text rejected; csg_last_error() carries the reference's message */
    CSG_ERR_IO = 2,        /* file could not be read */
    CSG_ERR_CUDA = 3,      /* a CUDA call failed; message = cudaGetErrorString */
    CSG_ERR_ARG = 4,       /* bad argument */
    CSG_ERR_NO_DEVICE = 5, /* no usable CUDA device (there is no CPU fallback) */
    CSG_ERR_LIMIT = 6      /* tree too deep / too large for the kernel's bounded stack */
};

/* Mirrors `Camera` (RenderManager/Camera/Camera.h:9-16): same field order, 60 bytes.
 * forward/right/up are the cached basis Camera::updateVectors computes (Camera.cpp:4-27);
 * fill them with csg_camera_set(), or copy them from an existing reference Camera. */
typedef struct csg_camera {
    float pos[3];
    float pitch, yaw; /* rotX, rotY in radians */
    float fov;        /* radians */
    float forward[3], right[3], up[3];
} csg_camera;

/* Mirrors `DirectionalLight` (RenderManager/DirectionalLight.h:8-18). */
typedef struct csg_light {
    float polar, azimuth; /* radians */
} csg_light;

/* ---- scene: replaces CSGTree::Parse (RayCasting/CSGTree/CSGTree.cuh:53, CSGTree.cu:5-152) and
 *      Application::LoadCSGTree (Application.cpp:59-83).  Grammar, keywords, argument counts,
 *      rotation range checks and error texts are the reference's; truncated input (undefined
 *      behaviour in the reference) is reported as CSG_ERR_PARSE. */
int csg_load_scene(const char* path, csg_scene** out);
int csg_parse_scene(const char* text, size_t len, csg_scene** out);
void csg_free_scene(csg_scene* scene);
/* n_nodes / n_prims as CSGTree::nodes.size() / primitives.size(); depth = operator levels on the longest path */
int csg_scene_counts(const csg_scene* scene, int* n_nodes, int* n_prims, int* depth);
/* Copies the tree in the reference's own layouts: n_nodes*44 bytes of CSGNode (CSGTree.cuh:15-34,
 * with the reference's AABBs of BVHNode.cuh:17-77) and n_prims*48 bytes of Primitive (Primitives.h:29-41). */
int csg_scene_dump(const csg_scene* scene, void* nodes44, void* prims48);
/* The GPU layout csg_upload builds from the scene at its current optimisation level (host side only, no device needed;
 * for tests and tools): *n_nodes 32-byte records in preorder (layout: cuda-csg-tree-raycasting_b200/csrc/csg_scene.h),
 * the parent of every record (-1 at the root) and the operator depth.  nodes32 / parents may be NULL to query the count. */
int csg_scene_flatten(const csg_scene* scene, void* nodes32, int32_t* parents, int* n_nodes, int* depth);
/* Writes the scene back out in the text format. Returns bytes needed (excluding NUL). */
size_t csg_scene_write(const csg_scene* scene, char* buf, size_t buflen);
/* Synthetic balanced tree in the scene text format (BASELINE.json configs[4]).
 * Returns bytes needed (excluding NUL); call with buf=NULL to size. */
size_t csg_generate_scene(int n_primitives, uint64_t seed, char* buf, size_t buflen);

/* ---- camera / light: replace Camera (Camera.h:18-52, Camera.cpp) and DirectionalLight */
void csg_camera_default(csg_camera* cam);                                        /* Camera(): pos (0,0,5), fov 90*3.14159/180 */
void csg_camera_set(csg_camera* cam, float x, float y, float z, float pitch, float yaw); /* setPosition + setRotation */
void csg_camera_set_fov_degrees(csg_camera* cam, float degrees);                 /* setFOV */
void csg_light_default(csg_light* light);                                        /* polar -60 deg, azimuth -45 deg */
void csg_light_direction(const csg_light* light, float out3[3]);                 /* getLightDir */

/* ---- upload: replaces Raycaster::ChangeSize(int w,int h,CSGTree) (RayCasting/Raycaster.cu:3-21).
 * In-process form: the frame is sharded in 64x32-pixel tiles over devices 0..n_gpus-1, every shard
 * writes its tiles straight into device 0's framebuffer over NVLink (peer stores). */
int csg_upload(const csg_scene* scene, int width, int height, int n_gpus, csg_context** out);
/* One-process-per-GPU form (torchrun): this process renders shard `shard_rank` of `shard_count` on
 * CUDA device `device`.  The gather target is set with csg_set_gather_target(). */
int csg_upload_shard(const csg_scene* scene, int width, int height, int device, int shard_rank,
                     int shard_count, csg_context** out);
void csg_free_context(csg_context* ctx); /* replaces Raycaster::CleanUp (Raycaster.cu:36-45) */

/* Diagnostic (host only): the cube normal of the reference, (float)(int)(((p - centre) / half_size) * 1.00001f) per component
 * (RaycastingKernels.cu:422-424), is a step function of |p - centre|; this returns the smallest value for which it reaches
 * `level` (1 or 2), as csg_upload stores it per cube.  0: half_size is not a positive normal number (no shortcut taken). */
float csg_cube_normal_threshold(float half_size, float level);

/* Load-time tree optimisation.  0 = keep the parsed tree shape; 1 (default) =
 * spatially re-balance maximal Union-only subtrees and tighten culling boxes.  Results are
 * identical except on exact-tie pixels; must be called before csg_upload. */
int csg_scene_set_optimize(csg_scene* scene, int level);

/* ---- render: replaces Raycaster::Raycast(float4* devPBO, Camera, DirectionalLight) (Raycaster.cu:23-34).
 * All three are synchronous (return when the output is complete), like the reference. */
/* rgba8_out: width*height*4 bytes; host OR device pointer (detected with cudaPointerGetAttributes).
 * Host pointer: the frame is dealt out to the context's GPUs in rows of 64x32-pixel tiles; every GPU renders its rows into its
 * own memory and copies them to rgba8_out over its own PCIe link, band by band while the next band renders (6 bands at 4K on
 * one GPU; use pinned memory — cudaHostAllocPortable / cudaHostRegisterPortable for several GPUs — for the copies to overlap).
 * On a csg_upload_shard context (one process per GPU) the call renders and copies THIS rank's rows only: give every rank the
 * same buffer (shared memory) and the frame is complete when every rank has returned.
 * Device pointer: multi-GPU contexts gather the frame in GPU 0's framebuffer over NVLink, then copy it. */
int csg_render(csg_context* ctx, const csg_camera* cam, const csg_light* light, uint8_t* rgba8_out);
/* rgba_f32_out: width*height float4, linear colour exactly as the reference's LightningKernel writes its
 * PBO (RaycastingKernels.cu:49-111); host or device pointer.  This is the literal PBO drop-in. */
int csg_render_f32(csg_context* ctx, const csg_camera* cam, const csg_light* light, float* rgba_f32_out);
/* Parity / debug outputs = RayHit.hit, RayHit.primitiveIdx (-1 on miss), RayHit.t (-1 on miss)
 * (RayCasting/Utils/Ray.cuh:24-34).  Host pointers; any may be NULL. */
int csg_render_aov(csg_context* ctx, const csg_camera* cam, uint8_t* hit, int32_t* prim_id, float* t);

/* Camera batch (BASELINE.json configs[1]: a 64-frame orbit).  Renders n_frames frames, camera cams[k] -> frame k at
 * rgba8_out + k*width*height*4 (host or device pointer), all with the same light.  Frames alternate between two internal
 * frame slots (own stream, pruned trees and framebuffer each), so frame k+1 is pruned and started while frame k drains and
 * is copied out; returns when every frame is complete.  csg_last_frame_ms() then reports the whole batch.  Single-GPU
 * contexts only.  Every frame is identical to what csg_render() produces for its camera. */
int csg_render_batch(csg_context* ctx, const csg_camera* cams, int n_frames, const csg_light* light, uint8_t* rgba8_out);

/* Supersampling (BASELINE.json configs[4]: 16 rays/pixel = 4 per axis).  Sub-sample (sx,sy) of pixel (x,y) is the reference's
 * ray generation at virtual pixel (x*k+sx, y*k+sy) of a (width*k) x (height*k) frame; the k*k linear colours are box-filtered.
 * I.e. the result equals the reference kernel run at k times the resolution and averaged k x k.  Default 1. */
int csg_set_supersampling(csg_context* ctx, int samples_per_axis);

/* Work statistics of THIS implementation (the reference algorithm's counts come from the instrumented oracle): per pixel, the
 * traversal loop iterations packed as search_visits << 20 | frame_machine_visits << 10 | other_iterations (nearest-Enter
 * search visits, operator visits of the frame machine, Compute/Return/leaf-loop rounds).  Host pointer, width*height int32. */
int csg_render_stats(csg_context* ctx, const csg_camera* cam, int32_t* iterations);

/* Per-tile tree pruning (on by default): before each frame every 64x32-pixel tile gets its own copy of the tree holding only
 * the primitives its rays can reach (operators left with one operand collapse to it).  Results are identical with and
 * without it.  mode 0: every tile reads the whole tree; 1 (default): per-tile trees built with prefix sums over
 * the preorder layout (the flat pruning kernel; trees of up to 2048 nodes, larger ones use the walk); 2: per-tile trees built
 * by the tree-walking pruning kernel.
 * csg_prune_stats reports the last frame of shard 0: traced tiles, tiles no primitive reaches, tiles whose tree did not fit
 * its slot (they read the whole tree), and the total number of nodes over all pruned trees. */
int csg_set_pruning(csg_context* ctx, int mode);
/* View cache (off by default): the per-tile trees depend on the camera, the frame size and the sampling only — not on the
 * light.  With the cache on, a frame whose view equals the previous frame's reuses the trees instead of rebuilding them
 * (the reference application's static camera with a moving light: one kernel per frame instead of two). */
int csg_set_view_cache(csg_context* ctx, int enabled);
int csg_prune_stats(csg_context* ctx, int* traced_tiles, int* empty_tiles, int* fallback_tiles, long long* pruned_nodes);

/* ---- asynchronous / device-resident form (benchmarks, interop viewers, multi-process gather) */
/* Enqueues one frame on the context's stream(s); rgba8_dev NULL = the context's own framebuffer
 * (or the gather target).  Returns without waiting. */
int csg_render_enqueue(csg_context* ctx, const csg_camera* cam, const csg_light* light, uint8_t* rgba8_dev);
int csg_sync(csg_context* ctx);
/* CUDA-event time (ms) of the last completed csg_render_enqueue / csg_render* call: first kernel start to
 * framebuffer complete on the root device. */
int csg_last_frame_ms(csg_context* ctx, float* ms);
/* The CUDA stream (a cudaStream_t, on the context's root device) that csg_render_enqueue puts the root GPU's share of a frame on.
 * Work a caller queues there itself — unmapping a GL buffer as RenderManager.cpp:58-84 does around Raycast, a post-process, a
 * benchmark's cache flush — is ordered against the frames on the device, without a host round trip; with work queued ahead of
 * it a frame's launches are already waiting when the GPU gets to them (the steady state of a render loop). */
int csg_stream(csg_context* ctx, void** cuda_stream);
/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
uint64_t csg_launch_count(const csg_context* ctx);
/* Device pointer of the context's own RGBA8 framebuffer (valid until csg_free_context). */
int csg_framebuffer(csg_context* ctx, uint8_t** rgba8_dev);
/* Multi-process gather: 64-byte cudaIpcMemHandle_t of this context's framebuffer (root rank) ... */
int csg_framebuffer_ipc_handle(csg_context* ctx, void* handle64);
/* ... and on the other ranks: open the root's handle and make it the target of csg_render_enqueue(NULL).  The handle also
 * carries the root's sync words: a sharded frame starts, on every GPU, when the root GPU starts it, and the root's frame is
 * complete (csg_sync / csg_last_frame_ms on rank 0) only when every rank's pixels have landed — all on the device, no host
 * round trip.  Every rank must enqueue every frame; a rank that waits more than 2 s for another fails with CSG_ERR_CUDA. */
int csg_set_gather_target_ipc(csg_context* ctx, const void* handle64);
/* Same inside one process (e.g. one host thread per GPU, each with its own csg_upload_shard context): `root` is the rank-0
 * context of the same frame; its framebuffer and sync words become this context's gather target.  (Shards of one frame on
 * the SAME device — a testing set-up — must enqueue the root first: a peer's kernels wait on the device for the root to start
 * and a full grid of them leaves the root no room.) */
int csg_set_gather_root(csg_context* ctx, csg_context* root);
/* Same, for a pointer that is already addressable from this context's device.  On a csg_upload_shard context such frames are
 * neither gated nor joined on the device (a bare pointer carries no sync words): the caller synchronises the ranks. */
int csg_set_gather_target(csg_context* ctx, uint8_t* rgba8_dev);

/* Page-locks a host buffer for every GPU of the process (cudaHostRegisterPortable), so that csg_render's device -> host
 * copies run at PCIe speed and overlap with rendering; works on memory the caller allocated any way it likes, including a
 * shared-memory mapping that several one-process-per-GPU ranks render into.  Undo with csg_unpin_host_buffer before freeing. */
int csg_pin_host_buffer(void* host, size_t bytes);
int csg_unpin_host_buffer(void* host);

/* Copies `bytes` from the framebuffer to a host buffer (synchronous). */
int csg_read_framebuffer(csg_context* ctx, uint8_t* rgba8_host);

/* tanf(fov/2) as the CUDA device evaluates it (RaycastingKernels.cu:15-16 runs tan() on the device); lets a host-side
 * checker reproduce ray generation bit for bit. */
int csg_device_tan_half_fov(csg_context* ctx, float fov, float* out);

/* Measured FP32 roofline of `device`: an FFMA-only kernel (8 independent chains per thread), best of 4, in TFLOP/s. */
int csg_fp32_peak_tflops(int device, float* tflops);

/* Description of the launch configuration chosen at upload, as JSON (threads, CTAs, smem, tree bytes...). */
const char* csg_context_info(csg_context* ctx);

/* Host-side view of the tile hand-out the kernels use (no device needed; for tests): a frame of macro_x x macro_y macro tiles
 * whose traced rectangle is (rm_x0, rm_y0, rm_w, rm_h), dealt out to shard_count shards by interleaved tiles (shard_mode 0) or
 * interleaved macro-tile rows (shard_mode 1).  Reports how many tiles shard_rank gets (*n_tiles), how many tree slots a shard
 * owns (*n_slots), and for its tile number `tile` the macro tile (*mx, *my) and the slot its tree is written to. */
int csg_shard_tile(int macro_x, int macro_y, int rm_x0, int rm_y0, int rm_w, int rm_h, int shard_mode, int shard_rank, int shard_count,
                   int tile, int* mx, int* my, int* slot, int* n_tiles, int* n_slots);

const char* csg_last_error(void);
const char* csg_version(void);

#ifdef __cplusplus
}
#endif
#endif
