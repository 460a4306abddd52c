// csg_render.cu — launcher and the C ABI of libcsg_b200 (include/csg_b200.h); the kernels are in csg_frame.cuh / csg_prune.cuh.
//
// Replaces Raycaster::{ChangeSize,Raycast,CleanUp} (RayCasting/Raycaster.cu:3-45) and the two kernels it
// launches (RayCasting/Kernels/RaycastingKernels.cu).  sm_100a only; there is no CPU path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/csg_b200.h"
#include "csg_kernel.cuh"
#include "csg_frame.cuh"
#include "csg_prune.cuh"
#include "csg_scene.h"

using namespace csgb;

// =========================================================================================== device code
namespace csgb {

// FP32 roofline probe: 8 independent FFMA chains per thread, nothing else.
__global__ void __launch_bounds__(256) csg_ffma_probe_kernel(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace csgb

// =========================================================================================== host side
struct csg_scene {
    Scene scene;
};

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
#define CU(call)                                                                     \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) return fail(CSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// inside create_context: a failing CUDA call frees what the half-built context already owns
#define CUC(call)                                                                    \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            const std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_); \
            csg_free_context(c);                                                     \
            return fail(CSG_ERR_CUDA, m_);                                           \
        }                                                                            \
    } while (0)

// Hit words keep the primitive id in 22 bits and node records their right operand in 24 (csg_kernel.cuh, csg_scene.h); the host
// builders recurse once per tree level.  Anything beyond is refused, not aliased / overflowed.
constexpr size_t kMaxPrims = (size_t)1 << 22, kMaxNodes = (size_t)1 << 24;
constexpr int kMaxParsedDepth = 16384;
int scene_within_limits(const Scene& sc)
{
    if (sc.prims.size() >= kMaxPrims || sc.nodes.size() >= kMaxNodes)
        return fail(CSG_ERR_LIMIT, "scene too large: " + std::to_string(sc.prims.size()) + " primitives / " + std::to_string(sc.nodes.size()) +
                                       " nodes (limits: 2^22 primitives, 2^24 nodes)");
    const int d = sc.depth();   // iterative
    if (d > kMaxParsedDepth)
        return fail(CSG_ERR_LIMIT, "tree depth " + std::to_string(d) + " exceeds " + std::to_string(kMaxParsedDepth) + " operator levels");
    return CSG_OK;
}

struct Shard {  // one GPU's share of the frame
    int device = 0;
    int rank = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr;
    PruneParams last_q;          // what the tile pool currently holds (view cache)
    bool last_q_valid = false;
    uint4* d_nodes = nullptr;
    uint4* d_pool = nullptr;     // [staged whole tree][one slot of slot_nodes records per macro tile of this shard]
    TileDesc* d_desc = nullptr;  // per macro tile of this shard
    int* d_parent = nullptr;
    uint2* d_topo = nullptr;             // per node: meta word, end of its subtree (csg_prune_flat_kernel)
    float4* d_leaf_boxes = nullptr;      // per primitive: culling box (world space) + node number, 2 x float4
    unsigned int* d_hist = nullptr;      // 2 x kCostBuckets counters: the bucket sizes of the last pruned frame, and the (zeroed) set of the next one
    int hist_cur = 0;                    // which set the tile lists currently belong to
    uint4* d_lists = nullptr;            // kCostBuckets x n_slots tile descriptors
    int n_slots = 0;
    float4* d_prims = nullptr;
    unsigned int* d_counter = nullptr;
    unsigned int counter_base = 0;
    float* d_tan = nullptr;              // device alias of h_tan
    float* h_tan = nullptr;              // page-locked, mapped: csg_tan_kernel writes tanf(fov/2) straight into host memory
    int grid = 0;
    int n_local_warp_tiles = 0;
    uint8_t* target = nullptr;   // where this shard writes RGBA8 (root framebuffer, possibly a peer pointer)
    void* ipc_mapped = nullptr;
    uint8_t* local_fb = nullptr;         // this shard's own full-size framebuffer (row-sharded frames for host output); shard 0: the context's
    bool owns_local_fb = false;
    cudaStream_t copy_stream = nullptr;  // device -> host copies of this shard's bands
    cudaEvent_t ev_band[8] = {};
    SyncWords* sync_words = nullptr;     // the ROOT's sync words as seen from this shard's device (nullptr: no gate / join)
    unsigned int* d_exit = nullptr;      // finished CTAs of the frame kernel (join)
    int* h_err = nullptr;                // mapped host word raised by a device-side wait that timed out
    int* d_err = nullptr;
};

}  // namespace

struct csg_context {
    int width = 0, height = 0;
    int macro_x = 0, macro_y = 0;
    int shard_count = 1;
    bool multi_process = false;
    FlatTree tree;
    bool has_cyl = true;         // the scene holds at least one cylinder: frame kernels with cylinder code (kCyl)
    bool prune = true;           // per-tile pruned trees, rebuilt every frame
    bool prune_flat = true;      // by csg_prune_flat_kernel (prefix sums over the preorder layout); false: csg_prune_kernel (tree walk)
    bool flat_ok = false;        // the tree is small enough for csg_prune_flat_kernel
    size_t flat_smem = 0;
    int slot_nodes = 0;          // records per tile slot
    size_t prune_smem = 0;
    uint32_t full_flags = 0;
    int mark_words = 0, marks_first = 0;
    csg_scene scene_copy;        // for the second frame slot of csg_render_batch
    csg_context* twin = nullptr; // second frame slot (own stream, trees, framebuffer), created on first use
    cudaEvent_t ev_batch0 = nullptr, ev_batch1 = nullptr;
    int warp_tree_nodes = 0;     // per-warp shared-memory copy of the current tile's tree: capacity in records
    int band_m0 = 0, band_m1 = 0;   // macro-tile rows the next enqueue covers (0, 0 = the whole frame)
    unsigned int frame_seq = 0;  // sequence number of the last sharded frame (gate / join words)
    size_t sync_off = 0;         // byte offset of the SyncWords behind the framebuffer (same allocation: one IPC handle covers both)
    int last_mode = 0;           // shard mode of the last frame (csg_prune_stats)
    bool shard_sync = true;      // sharded frames are started and joined on the device (SyncWords); csg_set_gather_target(pointer) turns it off
    bool view_cache = false;     // csg_set_view_cache
    int flat_leaves = kFlatLeavesMax;   // Unions over at most this many spheres are evaluated flat (eval_flat_union); 0: never
    bool external_target = false;   // csg_set_gather_target: pixels go to a buffer that is not rank 0's own framebuffer
    bool prune_alloc = false;    // tile slots were allocated at upload
    int last_rm[4] = {0, 0, 0, 0};   // traced macro-tile rectangle of the last frame (x0, y0, w, h)
    size_t smem_bytes = 0;
    int stack_levels = 1;
    int threads = 256;           // CTA shape chosen at upload (one of kShapeThreads)
    bool root_box_valid = false; // root_box = culling box of the whole scene (min xyz, max xyz), for the screen-space bound
    float root_box[6] = {0, 0, 0, 0, 0, 0};
    std::vector<Shard> shards;   // in-process: one per device; multi-process: exactly one
    uint8_t* d_fb = nullptr;     // RGBA8 framebuffer on the root device (or this rank's device)
    float* d_f32 = nullptr;      // lazily allocated
    uint8_t* d_aov_hit = nullptr;
    int32_t* d_aov_prim = nullptr;
    float* d_aov_t = nullptr;
    int32_t* d_aov_iters = nullptr;
    int ss = 1;                  // supersampling: samples per axis
    uint64_t launches = 0;
    float cached_fov = -1.f, cached_tan = 0.f;   // device tanf(fov/2) of the last field of view seen
    float last_ms = 0.f;
    bool frame_pending = false;
    std::string info;
};

namespace {

template <int MODE, int T>
int launch_one(csg_context* c, Shard& s, const FrameParams& fp)
{
    // Programmatic dependent launch: the frame kernel may start while csg_prune_kernel is still draining; it fills the
    // background macro tiles first and waits for the pruned trees (cudaGridDependencySynchronize) before tracing.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)s.grid);
    cfg.blockDim = dim3((unsigned)T);
    cfg.dynamicSmemBytes = c->smem_bytes;
    cfg.stream = s.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    // kernels without cylinder code for scenes without cylinders (instruction-cache footprint, csg_kernel.cuh eval_flat_union);
    // one ray per pixel: tickets of two warp tiles (kPair) when fill_params found plenty of them
    const bool pair = fp.pair_shift != 0;
    auto go = [&](auto kernel) { return cudaLaunchKernelEx(&cfg, kernel, fp); };
    if constexpr (MODE == OUT_AOV) {   // per primary ray: enqueue_frame refuses ss > 1
        e = c->has_cyl ? (pair ? go(csg_frame_kernel<MODE, T, false, true, true>) : go(csg_frame_kernel<MODE, T, false, true, false>))
                       : (pair ? go(csg_frame_kernel<MODE, T, false, false, true>) : go(csg_frame_kernel<MODE, T, false, false, false>));
    } else if (c->ss > 1) {
        e = c->has_cyl ? go(csg_frame_kernel<MODE, T, true, true, false>) : go(csg_frame_kernel<MODE, T, true, false, false>);
    } else {
        e = c->has_cyl ? (pair ? go(csg_frame_kernel<MODE, T, false, true, true>) : go(csg_frame_kernel<MODE, T, false, true, false>))
                       : (pair ? go(csg_frame_kernel<MODE, T, false, false, true>) : go(csg_frame_kernel<MODE, T, false, false, false>));
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CSG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return CSG_OK;
}

template <int MODE>
int launch_mode(csg_context* c, Shard& s, const FrameParams& fp)
{
    switch (c->threads) {
        case 768: return launch_one<MODE, 768>(c, s, fp);
        case 384: return launch_one<MODE, 384>(c, s, fp);
        default: return launch_one<MODE, 256>(c, s, fp);
    }
}

template <int MODE, int T>
int configure_one(size_t smem, int* blocks_per_sm)
{
    const int sm = (int)smem;
    if constexpr (MODE != OUT_AOV) {
        CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    }
    CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, csg_frame_kernel<MODE, T, false, true, false>, T, smem));
    return CSG_OK;
}

template <int T>
int configure_modes(size_t smem, int* bps)
{
    int rc;
    if ((rc = configure_one<OUT_F32, T>(smem, bps))) return rc;
    if ((rc = configure_one<OUT_AOV, T>(smem, bps))) return rc;
    return configure_one<OUT_RGBA8, T>(smem, bps);
}

int configure_shape(int threads, size_t smem, int* bps)
{
    switch (threads) {
        case 768: return configure_modes<768>(smem, bps);
        case 384: return configure_modes<384>(smem, bps);
        default: return configure_modes<256>(smem, bps);
    }
}

// Pixel rectangle that contains the projection of the root's culling box (padded by 2 pixels); full frame when any corner
// of the box is not safely in front of the camera.
void screen_bound(const csg_context* c, const csg_camera* cam, FrameParams& fp)
{
    fp.rect_x0 = 0; fp.rect_y0 = 0; fp.rect_x1 = c->width - 1; fp.rect_y1 = c->height - 1;
    if (!c->root_box_valid) return;
    const double W = (double)c->width * c->ss, H = (double)c->height * c->ss;
    const double th = std::tan((double)cam->fov * 0.5), aspect = W / H;
    if (!(th > 1e-6) || !std::isfinite(th)) return;
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int k = 0; k < 8; ++k) {
        double v[3];
        for (int a = 0; a < 3; ++a) v[a] = (double)((k >> a) & 1 ? c->root_box[3 + a] : c->root_box[a]) - (double)cam->pos[a];
        const double z = v[0] * cam->forward[0] + v[1] * cam->forward[1] + v[2] * cam->forward[2];
        const double len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (!(z > 1e-3 * len) || !(z > 1e-9)) return;   // corner beside/behind the camera: no useful bound
        const double nx = (v[0] * cam->right[0] + v[1] * cam->right[1] + v[2] * cam->right[2]) / z;
        const double ny = (v[0] * cam->up[0] + v[1] * cam->up[1] + v[2] * cam->up[2]) / z;
        const double px = (nx / (aspect * th) + 1.0) * 0.5 * (W - 1.0) - 0.5;    // inverse of RaycastKernel :11-16
        const double py = (1.0 - ny / th) * 0.5 * (H - 1.0) - 0.5;
        if (!std::isfinite(px) || !std::isfinite(py)) return;
        xmin = std::min(xmin, px); xmax = std::max(xmax, px);
        ymin = std::min(ymin, py); ymax = std::max(ymax, py);
    }
    const double pad = 2.0 * c->ss, big = 1e9;
    xmin = std::max(-big, xmin - pad); ymin = std::max(-big, ymin - pad);
    xmax = std::min(big, xmax + pad);  ymax = std::min(big, ymax + pad);
    fp.rect_x0 = (int)std::floor(xmin / c->ss); fp.rect_y0 = (int)std::floor(ymin / c->ss);
    fp.rect_x1 = (int)std::ceil(xmax / c->ss);  fp.rect_y1 = (int)std::ceil(ymax / c->ss);
}

void fill_params(const csg_context* c, const Shard& s, const csg_camera* cam, const float light[3], int shard_mode, FrameParams& fp)
{
    std::memset(&fp, 0, sizeof fp);
    for (int i = 0; i < 3; ++i) {
        fp.cam_pos[i] = cam->pos[i];
        fp.forward[i] = cam->forward[i];
        fp.right[i] = cam->right[i];
        fp.up[i] = cam->up[i];
        fp.light[i] = light ? light[i] : 0.f;
    }
    fp.fov = cam->fov;
    fp.tan_half_fov = c->cached_tan;
    fp.width = c->width;
    fp.height = c->height;
    fp.macro_x = c->macro_x;
    fp.macro_y = c->macro_y;
    // floor(n / d) == umulhi(n, floor(2^32 / d) + 1) whenever n < 2^20 and 1 < d < 2^12
    fp.div_magic = (c->macro_x > 1 && c->macro_x < 4096 && (long long)c->macro_x * c->macro_y < (1ll << 20))
                       ? (unsigned int)((1ull << 32) / (unsigned long long)c->macro_x + 1ull) : 0u;
    fp.shard_rank = s.rank;
    fp.shard_count = c->shard_count;
    fp.shard_mode = shard_mode;
    if (shard_mode) {
        // rows: every shard fills the background tiles of its own macro-tile rows (the kernel skips the other rows)
        fp.fill_stride = 1;
        fp.fill_first = 0;
    } else {
        // background tiles: the shard that owns the framebuffer fills all of them; with an external target everybody fills its own
        const bool owns_fb = s.rank == 0 && !c->external_target;
        fp.fill_stride = owns_fb ? 1 : c->shard_count;
        fp.fill_first = owns_fb ? 0 : (c->external_target ? s.rank : (1 << 30));
    }
    fp.counter_base = s.counter_base;
    fp.tile_counter = s.d_counter;
    fp.pool = s.d_pool;
    fp.desc = c->prune ? s.d_desc : nullptr;
    fp.lists = c->prune ? s.d_lists : nullptr;   // (hist: set by enqueue_frame once it knows whether this frame prunes)
    fp.list_stride = s.n_slots;
    fp.full_flags = c->full_flags;
    fp.prims = s.d_prims;
    fp.n_nodes = (int)c->tree.nodes.size();
    fp.root_is_leaf = c->tree.root_is_leaf ? 1 : 0;
    fp.root_pure = c->tree.root_pure ? 1 : 0;
    fp.stack_levels = c->stack_levels;
    fp.warp_tree_nodes = c->prune ? c->warp_tree_nodes : 0;
    fp.ss = c->ss;
    const float wf = (float)(c->width * c->ss), hf = (float)(c->height * c->ss);
    fp.wm1 = wf - 1.0f;       // (width - 1), :11
    fp.hm1 = hf - 1.0f;       // (height - 1), :12
    fp.aspect = wf / hf;      // (width / height), :15
    screen_bound(c, cam, fp);
    // macro-tile rectangle touched by the bound; the ticket space of this frame covers only these
    // a frame may be rendered in horizontal bands of macro-tile rows (csg_render to host memory: a band is copied out while
    // the next one renders); this launch covers rows [band_m0, band_m1)
    fp.band_m0 = c->band_m0;
    fp.band_m1 = c->band_m1 > 0 ? c->band_m1 : c->macro_y;
    const int ax0 = std::max(fp.rect_x0, 0), ay0 = std::max(std::max(fp.rect_y0, 0), fp.band_m0 * kMacroH);
    const int ax1 = std::min(fp.rect_x1, c->width - 1), ay1 = std::min(std::min(fp.rect_y1, c->height - 1), fp.band_m1 * kMacroH - 1);
    if (ax0 > ax1 || ay0 > ay1) {
        fp.rm_x0 = fp.rm_y0 = 0; fp.rm_w = fp.rm_h = 0;
    } else {
        fp.rm_x0 = ax0 / kMacroW; fp.rm_y0 = ay0 / kMacroH;
        fp.rm_w = ax1 / kMacroW - fp.rm_x0 + 1; fp.rm_h = ay1 / kMacroH - fp.rm_y0 + 1;
    }
    if (!shard_mode && c->shard_count > 1 && fp.rm_h > 1) {
        // Tiles are dealt out along the rows of the traced rectangle, tile j to shard j mod N.  When the rectangle's width shares a
        // factor with N, a shard's tiles line up in columns (width 24 on 8 GPUs: shard r owns columns r, r + 8, r + 16 of the frame)
        // and the shards' loads differ by what those columns hold — Cheese512 @ 4K on 8 GPUs: the heaviest shard carried 8.5 % more
        // than the mean and finished 6 us after the others.  One more column of tiles (they lie outside the screen-space bound:
        // empty trees, background pixels, like the corners of the rectangle) makes the width coprime to N, so that every row is
        // shifted against the one above: max / mean 1.085 -> 1.016 on that frame (4 GPUs: 1.062 -> 1.007).
        for (int extra = 0; extra < 2 && std::gcd(fp.rm_w, c->shard_count) != 1; ++extra) {
            if (fp.rm_x0 + fp.rm_w < c->macro_x) ++fp.rm_w;
            else if (fp.rm_x0 > 0) { --fp.rm_x0; ++fp.rm_w; }
            else break;
        }
    }
    const long long traced = (long long)fp.rm_w * fp.rm_h;
    fp.row_first = shard_row_first(fp.rm_y0, s.rank, c->shard_count);
    const long long mine = shard_mode ? (long long)shard_row_count(fp.rm_y0, fp.rm_h, s.rank, c->shard_count) * fp.rm_w
                                      : (traced > s.rank ? (traced - s.rank + c->shard_count - 1) / c->shard_count : 0);
    fp.sp_shift = (c->ss == 2) ? 2 : (c->ss == 4) ? 4 : 0;   // sample-parallel supersampling: 4 or 16 rays per pixel
    {
        const char* serial = std::getenv("CSG_B200_SERIAL_SS");   // testing aid: loop over the samples in one lane
        if (serial && serial[0] == '1') fp.sp_shift = 0;
    }
    // passes per ticket: 1 << sp_group; default: a warp tile in two tickets (measured best on the 8K x 16 spp config at 1..8 GPUs)
    fp.sp_group = fp.sp_shift > 1 ? fp.sp_shift - 1 : 0;
    {
        const char* g = std::getenv("CSG_B200_SS_GROUP");   // tuning aid
        if (g) fp.sp_group = std::min(std::max(std::atoi(g), 0), fp.sp_shift);
    }
    fp.sp_tshift = fp.sp_shift - fp.sp_group;
    fp.n_local_warp_tiles = (int)((mine * 64) << fp.sp_tshift);
    fp.pair_shift = 0;
    if (c->ss == 1) {
        // one ray per pixel: tickets of two warp tiles while there are plenty of them (12 warp tiles per warp of the grid or more: one
        // GPU at 4K), single warp tiles when they are scarce (sharded frames)
        const long long warps = (long long)s.grid * (c->threads / 32);
        fp.pair_shift = mine * 64 >= 12 * warps ? 1 : 0;
        if (const char* e = std::getenv("CSG_B200_PAIR")) fp.pair_shift = std::atoi(e) ? 1 : 0;   // tuning aid
        fp.n_local_warp_tiles = (int)((mine * 64) >> fp.pair_shift);
    }
    fp.rm_magic = (fp.rm_w > 1 && fp.rm_w < 4096 && traced < (1ll << 20)) ? (unsigned int)((1ull << 32) / (unsigned long long)fp.rm_w + 1ull) : 0u;
    if (fp.rm_w == 0) fp.rm_w = 1;   // never divide by zero; n_local_warp_tiles is 0 anyway
}

// tanf(fov / 2) as the device evaluates it.  The kernel writes into mapped page-locked host memory and only the context's own
// stream is waited for: no copy engine and no device-wide synchronisation are involved, so this also works while another
// shard's kernels of the same frame sit on the device waiting for this shard to start (the gate of sharded frames).
int device_tan(csg_context* c, float fov, float* out)
{
    Shard& root = c->shards[0];
    CU(cudaSetDevice(root.device));
    csg_tan_kernel<<<1, 1, 0, root.stream>>>(fov, root.d_tan);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(root.stream));
    *out = *reinterpret_cast<volatile float*>(root.h_tan);
    c->launches++;
    return CSG_OK;
}

// Enqueue one frame on every shard.  mode: OUT_RGBA8 -> out = rgba8 target (NULL: each shard's own target),
// OUT_F32 / OUT_AOV only on single-shard contexts.
// shard_mode 0: macro tiles interleaved over the shards, every shard stores into one framebuffer (the root's, over NVLink), the
//               frame is started and joined on the device through the root's SyncWords (GateParams);
// shard_mode 1: macro-tile rows interleaved over the shards, every shard renders into its own local framebuffer (host output:
//               each shard then copies its own rows over its own PCIe link); no dependency between the shards.
int enqueue_frame(csg_context* c, const csg_camera* cam, const float light[3], int mode, void* out, int shard_mode = 0)
{
    if (!c || !cam) return fail(CSG_ERR_ARG, "null argument");
    if (mode != OUT_RGBA8 && c->shards.size() != 1) return fail(CSG_ERR_ARG, "f32/aov output needs a single-shard context");
    if (mode == OUT_AOV && c->ss != 1) return fail(CSG_ERR_ARG, "AOV output is per primary ray: set supersampling to 1");
    Shard& root = c->shards[0];
    CU(cudaSetDevice(root.device));
    if (!(cam->fov == c->cached_fov)) {   // new field of view: one tiny launch, then cached
        int rc = device_tan(c, cam->fov, &c->cached_tan);
        if (rc) return rc;
        c->cached_fov = cam->fov;
    }
    // device-side start gate + join: frames whose shards all store into the root's framebuffer
    const bool joined = c->shard_count > 1 && shard_mode == 0 && mode == OUT_RGBA8 && c->shard_sync;
    if (joined) ++c->frame_seq;
    c->last_mode = shard_mode;
    const int total_warps_per_cta = c->threads / 32;
    // The other shards first, the root last: a peer's kernels wait on the device for the root's start word, so by the time
    // the root's first kernel runs (right behind its start event) the peers of an in-process context are already queued.
    for (size_t k = c->shards.size(); k-- > 0;) {
        Shard& s = c->shards[k];
        CU(cudaSetDevice(s.device));
        FrameParams fp;
        fill_params(c, s, cam, light, shard_mode, fp);
        if (&s == &root) { c->last_rm[0] = fp.rm_x0; c->last_rm[1] = fp.rm_y0; c->last_rm[2] = fp.rm_h ? fp.rm_w : 0; c->last_rm[3] = fp.rm_h; }
        GateParams gate;
        std::memset(&gate, 0, sizeof gate);
        if (joined && s.sync_words) {
            gate.role = s.rank == 0 ? GATE_ROOT : GATE_PEER;
            gate.seq = c->frame_seq;
            gate.n_shards = c->shard_count;
            gate.rank = s.rank;
            gate.words = s.sync_words;
            gate.exit_counter = s.d_exit;
            gate.err = s.d_err;
            gate.enter = 1;
            gate.local_start = s.d_exit + 32;   // (its own 128-byte line)
        }
        bool gate_entered = false;   // a kernel of this frame in front of the frame kernel has been through the gate
        {   // per-tile pruned trees + the staged copy of the whole tree for this camera
            PruneParams q;
            std::memset(&q, 0, sizeof q);
            for (int i = 0; i < 3; ++i) {
                q.cam_pos[i] = fp.cam_pos[i]; q.forward[i] = fp.forward[i]; q.right[i] = fp.right[i]; q.up[i] = fp.up[i];
            }
            q.tan_half_fov = fp.tan_half_fov; q.wm1 = fp.wm1; q.hm1 = fp.hm1; q.aspect = fp.aspect; q.ss = fp.ss;
            q.width = fp.width; q.height = fp.height; q.macro_x = fp.macro_x;
            q.rm_x0 = fp.rm_x0; q.rm_y0 = fp.rm_y0; q.rm_w = fp.rm_w; q.rm_magic = fp.rm_magic;
            q.shard_rank = fp.shard_rank; q.shard_count = fp.shard_count; q.shard_mode = fp.shard_mode; q.row_first = fp.row_first;
            q.n_tiles = c->prune ? (int)(((long long)fp.n_local_warp_tiles << fp.pair_shift) >> fp.sp_tshift) / 64 : 0;
            q.nodes = s.d_nodes; q.n_nodes = fp.n_nodes; q.parent = s.d_parent;
            q.leaf_boxes = s.d_leaf_boxes; q.n_leaves = (int)(c->tree.leaf_boxes.size() / 8);
            q.mark_words = c->mark_words; q.marks_first = c->marks_first;
            q.topo = s.d_topo;
            q.prims = s.d_prims; q.n_prims = (int)c->tree.prims.size();
            q.n_slots = s.n_slots; q.lists = s.d_lists;   // (hist / hist_next: below, they alternate and must stay out of the view-cache comparison)
            // (heaviest-first hand-out also pays on a frame sharded over 8 GPUs, 2-3 tiles per warp: the slowest of the eight shards
            // takes 63.9 us with it and 67.3 us with the natural order, although shard 0 alone is 3 us faster without the ordering pass)
            q.pool = s.d_pool; q.desc = s.d_desc; q.slot_nodes = c->slot_nodes; q.flat_max = c->flat_leaves; q.flat_tree_max = fp.warp_tree_nodes;
            q.slots_off32 = (uint32_t)fp.n_nodes; q.full_flags = c->full_flags;
            // view cache (opt-in): same camera, size, sampling and tile set as the trees this shard already holds -> keep them
            // (compared before the gate goes in: its sequence number changes with every frame)
            const bool cached = c->view_cache && s.last_q_valid && std::memcmp(&q, &s.last_q, sizeof q) == 0;
            s.last_q = q;
            s.last_q_valid = true;
            q.gate = gate;
            if (s.d_hist) {
                // a pruned frame counts into the set of counters the previous pruned frame zeroed, and zeroes the other one
                const int set = cached ? s.hist_cur : s.hist_cur ^ 1;
                q.hist = s.d_hist + set * kCostBuckets;
                q.hist_next = s.d_hist + (set ^ 1) * kCostBuckets;
                s.hist_cur = set;
                fp.hist = q.hist;
            }
            const int stage_ctas = std::max(1, std::min(64, (fp.n_nodes + kPruneThreads - 1) / kPruneThreads));
            // the frame's start mark goes in right in front of its first launch (all host-side preparation is done by now: on an
            // idle GPU whatever the host does between the two calls would show up as device time)
            if (&s == &root && c->band_m0 == 0) CU(cudaEventRecord(root.ev_start, root.stream));   // later bands of a banded frame keep the first band's start mark
            if (!cached) {
                if (c->prune_flat && c->flat_ok) {
                    // CTA size by the number of tiles: with few tiles per SM the per-tile latency is what the frame waits for
                    int sms = 148;
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device);
                    int ft = q.n_tiles <= sms ? 512 : q.n_tiles <= 3 * sms ? 256 : 128;
                    if (const char* e = std::getenv("CSG_B200_FLAT_THREADS")) ft = std::atoi(e);   // tuning aid
                    const unsigned grid = (unsigned)(q.n_tiles + stage_ctas);
                    if (ft == 512) csg_prune_flat_kernel<512><<<grid, 512, c->flat_smem, s.stream>>>(q);
                    else if (ft == 256) csg_prune_flat_kernel<256><<<grid, 256, c->flat_smem, s.stream>>>(q);
                    else csg_prune_flat_kernel<128><<<grid, 128, c->flat_smem, s.stream>>>(q);
                }
                else csg_prune_kernel<<<(q.n_tiles + kPruneWarps - 1) / kPruneWarps + stage_ctas, kPruneThreads, c->prune_smem, s.stream>>>(q);
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return fail(CSG_ERR_CUDA, std::string("prune kernel launch: ") + cudaGetErrorString(e));
            if (!cached) { c->launches++; gate_entered = true; }
        }
        fp.gate = gate;
        if (gate_entered) fp.gate.enter = 0;   // the frame kernel runs behind the pruning kernel (grid dependency), which has opened / passed the gate
        int rc = CSG_OK;
        if (mode == OUT_RGBA8) {
            fp.out = out ? out : (void*)(shard_mode ? s.local_fb : s.target);
            if (!fp.out) return fail(CSG_ERR_ARG, "no output target");
            rc = launch_mode<OUT_RGBA8>(c, s, fp);
        } else if (mode == OUT_F32) {
            fp.out = out;
            rc = launch_mode<OUT_F32>(c, s, fp);
        } else {
            fp.aov_hit = c->d_aov_hit;
            fp.aov_prim = c->d_aov_prim;
            fp.aov_t = c->d_aov_t;
            fp.aov_iters = c->d_aov_iters;
            rc = launch_mode<OUT_AOV>(c, s, fp);
        }
        if (rc) return rc;
        c->launches++;
        // every warp of the grid draws exactly one ticket past the end
        s.counter_base += (unsigned int)fp.n_local_warp_tiles + (unsigned int)(s.grid * total_warps_per_cta);   // wraps mod 2^32 like the device counter
        if (&s != &root) CU(cudaEventRecord(s.ev_done, s.stream));
    }
    CU(cudaSetDevice(root.device));
    if (!joined || !root.sync_words)   // no device-side join (row-sharded frames, or no sync words): the root's done event follows the peers' events
        for (size_t i = 1; i < c->shards.size(); ++i) CU(cudaStreamWaitEvent(root.stream, c->shards[i].ev_done, 0));
    CU(cudaEventRecord(root.ev_done, root.stream));
    c->frame_pending = true;
    return CSG_OK;
}

int sync_frame(csg_context* c)
{
    Shard& root = c->shards[0];
    CU(cudaSetDevice(root.device));
    CU(cudaEventSynchronize(root.ev_done));
    if (c->frame_pending) {
        CU(cudaEventElapsedTime(&c->last_ms, root.ev_start, root.ev_done));
        c->frame_pending = false;
    }
    CU(cudaGetLastError());
    for (Shard& s : c->shards)
        if (s.h_err && *reinterpret_cast<volatile int*>(s.h_err)) {
            *s.h_err = 0;
            return fail(CSG_ERR_CUDA, "sharded frame: a device-side wait for another shard timed out (did every rank enqueue the frame?)");
        }
    return CSG_OK;
}

bool is_device_pointer(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int create_context(const csg_scene* scene, int width, int height, const std::vector<int>& devices, int shard_rank0,
                   int shard_count, bool multi_process, csg_context** out)
{
    if (!scene || !out) return fail(CSG_ERR_ARG, "null argument");
    if (width < 2 || height < 2) return fail(CSG_ERR_ARG, "width and height must be >= 2");  // (w-1),(h-1) divisors, Q1
    if ((long long)width * height >= (1ll << 31)) return fail(CSG_ERR_ARG, "width*height must be below 2^31");
    // the kernels number macro tiles with a multiply-high division that is exact below 2^20 tiles of fewer than 4096 per row
    if ((width + kMacroW - 1) / kMacroW >= 4096 || (long long)((width + kMacroW - 1) / kMacroW) * ((height + kMacroH - 1) / kMacroH) >= (1ll << 20))
        return fail(CSG_ERR_LIMIT, "frame too large: at most 4095 x 64 pixels wide and 2^20 tiles of 64x32 pixels");
    if (scene->scene.nodes.empty()) return fail(CSG_ERR_ARG, "empty scene");
    if (int rc = scene_within_limits(scene->scene)) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CSG_ERR_NO_DEVICE, "no CUDA device available (libcsg_b200 has no CPU fallback)");
    }
    for (int d : devices)
        if (d < 0 || d >= ndev) return fail(CSG_ERR_NO_DEVICE, "requested CUDA device " + std::to_string(d) + " of " + std::to_string(ndev));

    csg_context* c = new csg_context();
    c->width = width;
    c->height = height;
    c->macro_x = (width + kMacroW - 1) / kMacroW;
    c->macro_y = (height + kMacroH - 1) / kMacroH;
    c->shard_count = shard_count;
    c->multi_process = multi_process;
    c->scene_copy.scene = scene->scene;
    flatten(scene->scene, scene->scene.optimize, c->tree);
    c->has_cyl = false;
    for (const NodeRec& nr : c->tree.nodes) c->has_cyl = c->has_cyl || (nr.meta & 7u) == (uint32_t)kCylinder;
    c->stack_levels = std::max(1, c->tree.depth);
    c->root_box_valid = c->tree.root_box_valid;
    for (int i = 0; i < 6; ++i) c->root_box[i] = c->tree.root_box[i];

    auto cleanup_fail = [&](int code) { const std::string keep = g_err; csg_free_context(c); g_err = keep; return code; };

    // shared memory plan: [table 128 B][stack 16 B x (levels+3) x threads][tree 32 B/node]
    CUC(cudaSetDevice(devices[0]));
    int max_optin = 0, sms = 0;
    CUC(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, devices[0]));
    CUC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devices[0]));
    const size_t tree_bytes = c->tree.nodes.size() * sizeof(NodeRec);
    const size_t table_bytes = kSmemHead;   // outcome table, light
    {
        // per-tile pruning: a slot holds up to kSlotMax records (8 KB)
        const size_t n = c->tree.nodes.size();
        c->full_flags = (c->tree.root_is_leaf ? kTileRootLeaf : 0u) | (c->tree.root_pure ? kTileRootPure : 0u);
        c->slot_nodes = (int)std::min<size_t>(n, kSlotMax);
        // leaf marks: 2 bits per node and warp; above 128K nodes (32 KB per warp) only the frustum walk is available
        c->mark_words = n <= 131072 ? (int)((n + 15) / 16) : 0;
        const char* mf = std::getenv("CSG_B200_MARKS_FIRST");   // tuning aid
        c->marks_first = mf ? (mf[0] == '1') : 0;   // the frustum walk first; the marks when it overflows
        c->prune_smem = kPruneWarps * (sizeof(PruneWarpSmem) + (size_t)c->mark_words * 4);
        const char* off = std::getenv("CSG_B200_NO_PRUNE");   // tuning aid
        // A scene that is a single primitive has nothing to prune, and its root is intersected without the gating box (Q7):
        // dropping a root cylinder by that (non-conservative, Q6) box would lose pixels the reference draws.
        c->prune = !(off && off[0] == '1') && !c->tree.root_is_leaf;
        c->prune_alloc = c->prune;
        // csg_prune_flat_kernel: two 16-bit prefix sums per node behind the fixed part of its shared memory
        // The flat kernel looks at every node of the tree for every tile; the walk only at what the frustum touches.  Measured
        // (B200): 1025 nodes x 1045 tiles 20 vs 34 us; 8191 nodes x 1800 tiles 570 vs 200 us -> flat up to kFlatPreferNodes.
        const char* force_flat = std::getenv("CSG_B200_PRUNE_FLAT");   // tuning aid: the flat kernel whenever it fits
        c->flat_ok = n <= (size_t)((force_flat && force_flat[0] == '1') ? kFlatMaxNodes : kFlatPreferNodes);
        c->flat_smem = sizeof(FlatTileSmem) + 2 * ((n + 7) & ~(size_t)7) * sizeof(unsigned short);
        if (const char* fl = std::getenv("CSG_B200_FLAT_LEAVES")) c->flat_leaves = std::min(std::max(std::atoi(fl), 0), kFlatLeavesMax);   // tuning aid
        const char* walk = std::getenv("CSG_B200_PRUNE_WALK");   // tuning aid: the tree-walking kernel instead
        c->prune_flat = !(walk && walk[0] == '1');
    }
    {
        // resident warps per SM for every (shape, tree placement); +1 KB per CTA is what the driver reserves
        const size_t sm_total = (size_t)max_optin + 1024;
        int best_warps = -1;
        const char* force_shape = std::getenv("CSG_B200_SHAPE");   // tuning aid: force a CTA shape
        for (int i = 0; i < kShapes; ++i) {
            const int T = kShapeThreads[i];
            if (force_shape && std::atoi(force_shape) != T) continue;
            const size_t base_need = (size_t)(c->stack_levels + 3) * T * sizeof(uint4) + table_bytes;   // +3: sentinel frame, search marker, scratch frame (supersampling)
            if (base_need > (size_t)max_optin) continue;
            const int ctas = (int)std::min<size_t>(min_blocks_for(T), sm_total / (base_need + 1024));
            const int warps = ctas * T / 32;
            if (warps > best_warps) {
                best_warps = warps; c->threads = T;
                // per-warp tree copies out of what is left of the SM's shared memory: 64, 32 or 0 nodes per warp
                c->warp_tree_nodes = 0;
                for (int cap : {64, 32}) {
                    const size_t need = base_need + (size_t)(T / 32) * cap * sizeof(NodeRec);
                    if (c->prune && need <= (size_t)max_optin && (size_t)ctas * (need + 1024) <= sm_total) { c->warp_tree_nodes = cap; break; }
                }
                c->smem_bytes = base_need + (size_t)(T / 32) * c->warp_tree_nodes * sizeof(NodeRec);
            }
        }
        if (best_warps <= 0) {
            g_err = "tree depth " + std::to_string(c->tree.depth) + " needs more traversal stack than one SM's shared memory (" +
                    std::to_string(max_optin) + " bytes) holds (the reference's own stacks hold 32 entries, RaycastingKernels.cuh:19)";
            return cleanup_fail(CSG_ERR_LIMIT);
        }
    }

    const int total_macros = c->macro_x * c->macro_y;
    c->shards.resize(devices.size());
    for (size_t i = 0; i < devices.size(); ++i) {
        Shard& s = c->shards[i];
        s.device = devices[i];
        s.rank = shard_rank0 + (int)i;
        CUC(cudaSetDevice(s.device));
        int bps = 0, rc;
        rc = configure_shape(c->threads, c->smem_bytes, &bps);
        if (rc) { const std::string keep = g_err; csg_free_context(c); g_err = keep; return rc; }
        if (bps < 1) { g_err = "kernel does not fit on an SM"; return cleanup_fail(CSG_ERR_LIMIT); }
        int dev_sms = 0;
        CUC(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, s.device));
        const int my_macros = slots_per_shard(c->macro_x, c->macro_y, shard_count);   // upper bound of this shard's macro tiles, either shard mode
        s.n_local_warp_tiles = my_macros * 64;
        const int want = (s.n_local_warp_tiles + (c->threads / 32) - 1) / (c->threads / 32);
        if (const char* cap = std::getenv("CSG_B200_CTAS_PER_SM")) bps = std::max(1, std::min(bps, std::atoi(cap)));   // tuning aid
        s.grid = std::max(1, std::min(dev_sms * bps, want));   // persistent CTAs: a multiple of the SM count
        CUC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CUC(cudaEventCreate(&s.ev_start));
        CUC(cudaEventCreate(&s.ev_done));
        CUC(cudaMalloc(&s.d_nodes, std::max<size_t>(tree_bytes, 32)));
        CUC(cudaMemcpy(s.d_nodes, c->tree.nodes.data(), tree_bytes, cudaMemcpyHostToDevice));
        {
            // slot of macro tile m = m / shard_count (csg_prune.cuh); a shard is handed tiles by their position inside the traced
            // rectangle, so any macro tile of the frame can come its way: ceil(total / count) slots, whatever the rank
            s.n_slots = slots_per_shard(c->macro_x, c->macro_y, shard_count);
            const size_t pool_records = c->tree.nodes.size() + (c->prune ? (size_t)s.n_slots * c->slot_nodes : 0);
            CUC(cudaMalloc(&s.d_pool, std::max<size_t>(pool_records, 1) * sizeof(NodeRec)));
            CUC(cudaMalloc(&s.d_desc, std::max<size_t>(s.n_slots, 1) * sizeof(TileDesc)));
            CUC(cudaMemset(s.d_desc, 0, std::max<size_t>(s.n_slots, 1) * sizeof(TileDesc)));
            CUC(cudaMalloc(&s.d_parent, std::max<size_t>(c->tree.parent.size(), 1) * sizeof(int)));
            CUC(cudaMemcpy(s.d_parent, c->tree.parent.data(), c->tree.parent.size() * sizeof(int), cudaMemcpyHostToDevice));
            CUC(cudaMalloc(&s.d_leaf_boxes, std::max<size_t>(c->tree.leaf_boxes.size(), 8) * sizeof(float)));
            CUC(cudaMemcpy(s.d_leaf_boxes, c->tree.leaf_boxes.data(), c->tree.leaf_boxes.size() * sizeof(float), cudaMemcpyHostToDevice));
            const char* no_order = std::getenv("CSG_B200_NO_ORDER");   // tuning aid: tiles handed out in their natural order
            if (c->prune && s.n_slots <= 65535 && !(no_order && no_order[0] == '1')) {   // bounds the bucket lists (64 x n_slots x 16 bytes)
                CUC(cudaMalloc(&s.d_hist, 2 * kCostBuckets * sizeof(unsigned int)));
                CUC(cudaMemset(s.d_hist, 0, 2 * kCostBuckets * sizeof(unsigned int)));
                CUC(cudaMalloc(&s.d_lists, (size_t)kCostBuckets * std::max(s.n_slots, 1) * sizeof(uint4)));
            }
            if (c->prune_smem > 48 * 1024)
                CUC(cudaFuncSetAttribute(csg_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->prune_smem));
            if (c->flat_ok) {
                std::vector<uint2> topo(c->tree.nodes.size());
                for (size_t k = 0; k < topo.size(); ++k) topo[k] = make_uint2(c->tree.nodes[k].meta, c->tree.subtree_end[k]);
                CUC(cudaMalloc(&s.d_topo, std::max<size_t>(topo.size(), 1) * sizeof(uint2)));
                CUC(cudaMemcpy(s.d_topo, topo.data(), topo.size() * sizeof(uint2), cudaMemcpyHostToDevice));
                if (c->flat_smem > 48 * 1024) {
                    CUC(cudaFuncSetAttribute(csg_prune_flat_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->flat_smem));
                    CUC(cudaFuncSetAttribute(csg_prune_flat_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->flat_smem));
                    CUC(cudaFuncSetAttribute(csg_prune_flat_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->flat_smem));
                }
            }
        }
        const size_t prim_bytes = c->tree.prims.size() * sizeof(PrimRec);
        CUC(cudaMalloc(&s.d_prims, std::max<size_t>(prim_bytes, 80)));
        CUC(cudaMemcpy(s.d_prims, c->tree.prims.data(), prim_bytes, cudaMemcpyHostToDevice));
        CUC(cudaHostAlloc(&s.h_tan, sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
        CUC(cudaHostGetDevicePointer(&s.d_tan, s.h_tan, 0));
        CUC(cudaMalloc(&s.d_counter, sizeof(unsigned int)));
        CUC(cudaMemset(s.d_counter, 0, sizeof(unsigned int)));
        CUC(cudaMalloc(&s.d_exit, 64 * sizeof(unsigned int)));   // [0]: finished CTAs; [32]: this shard's copy of the start word (GateParams::local_start)
        CUC(cudaMemset(s.d_exit, 0, 64 * sizeof(unsigned int)));
        CUC(cudaHostAlloc(&s.h_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        *s.h_err = 0;
        CUC(cudaHostGetDevicePointer(&s.d_err, s.h_err, 0));
        if (i == 0) {
            // the framebuffer, and behind it (same allocation, so that one IPC handle covers both) the SyncWords of sharded frames
            c->sync_off = (((size_t)width * height * 4) + 255) & ~(size_t)255;
            CUC(cudaMalloc(&c->d_fb, c->sync_off + sizeof(SyncWords)));
            CUC(cudaMemset(c->d_fb, 0, c->sync_off + sizeof(SyncWords)));
            s.local_fb = c->d_fb;
        } else {
            // NVLink peer stores into the root framebuffer
            int can = 0;
            CUC(cudaDeviceCanAccessPeer(&can, s.device, devices[0]));
            if (!can) { g_err = "device " + std::to_string(s.device) + " cannot access device " + std::to_string(devices[0]) + " (P2P)"; return cleanup_fail(CSG_ERR_CUDA); }
            cudaError_t pe = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { g_err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe); return cleanup_fail(CSG_ERR_CUDA); }
            cudaGetLastError();
        }
        s.target = c->d_fb;
        // in-process: every shard sees the root's words through peer access; one process per GPU: the root has them, the
        // others get them with the root's IPC handle (csg_set_gather_target_ipc)
        s.sync_words = (shard_count > 1 && (!multi_process || s.rank == 0)) ? reinterpret_cast<SyncWords*>(c->d_fb + c->sync_off) : nullptr;
    }
    char buf[512];
    std::snprintf(buf, sizeof buf,
                  "{\"threads_per_cta\": %d, \"ctas\": %d, \"sms\": %d, \"smem_bytes_per_cta\": %zu, \"tree_bytes\": %zu, "
                  "\"prune\": %s, \"prune_kernel\": \"%s\", \"slot_nodes\": %d, \"stack_levels\": %d, \"n_nodes\": %zu, \"n_prims\": %zu, \"shards\": %d, "
                  "\"macro_tiles\": %d, \"optimize\": %d}",
                  c->threads, c->shards[0].grid, sms, c->smem_bytes, tree_bytes, c->prune ? "true" : "false",
                  !c->prune ? "none" : (c->prune_flat && c->flat_ok) ? "prefix sums (csg_prune_flat_kernel)" : "tree walk (csg_prune_kernel)", c->slot_nodes,
                  c->stack_levels, c->tree.nodes.size(), c->tree.prims.size(), shard_count, total_macros, scene->scene.optimize);
    c->info = buf;
    *out = c;
    return CSG_OK;
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

const char* csg_last_error(void) { return g_err.c_str(); }
const char* csg_version(void) { return "csg_b200 0.1 (sm_100a)"; }

int csg_parse_scene(const char* text, size_t len, csg_scene** out)
{
    if (!text || !out) return fail(CSG_ERR_ARG, "null argument");
    csg_scene* s = new csg_scene();
    std::string err = parse_scene(text, len, s->scene);
    if (!err.empty()) {
        delete s;
        return fail(CSG_ERR_PARSE, err);
    }
    s->scene.optimize = 1;
    *out = s;
    return CSG_OK;
}

int csg_load_scene(const char* path, csg_scene** out)
{
    if (!path || !out) return fail(CSG_ERR_ARG, "null argument");
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(CSG_ERR_IO, std::string("cannot open ") + path);
    std::stringstream ss;
    ss << f.rdbuf();   // Application::LoadCSGTree reads the whole file, Application.cpp:59-75
    const std::string text = ss.str();
    return csg_parse_scene(text.data(), text.size(), out);
}

void csg_free_scene(csg_scene* scene) { delete scene; }

int csg_scene_counts(const csg_scene* scene, int* n_nodes, int* n_prims, int* depth)
{
    if (!scene) return fail(CSG_ERR_ARG, "null scene");
    if (n_nodes) *n_nodes = (int)scene->scene.nodes.size();
    if (n_prims) *n_prims = (int)scene->scene.prims.size();
    if (depth) *depth = scene->scene.depth();
    return CSG_OK;
}

int csg_scene_dump(const csg_scene* scene, void* nodes44, void* prims48)
{
    if (!scene) return fail(CSG_ERR_ARG, "null scene");
    if (nodes44) std::memcpy(nodes44, scene->scene.nodes.data(), scene->scene.nodes.size() * sizeof(RefNode));
    if (prims48) std::memcpy(prims48, scene->scene.prims.data(), scene->scene.prims.size() * sizeof(RefPrim));
    return CSG_OK;
}

int csg_scene_flatten(const csg_scene* scene, void* nodes32, int32_t* parents, int* n_nodes, int* depth)
{
    if (!scene) return fail(CSG_ERR_ARG, "null scene");
    if (int rc = scene_within_limits(scene->scene)) return rc;
    FlatTree t;
    flatten(scene->scene, scene->scene.optimize, t);
    if (nodes32) std::memcpy(nodes32, t.nodes.data(), t.nodes.size() * sizeof(NodeRec));
    if (parents) std::memcpy(parents, t.parent.data(), t.parent.size() * sizeof(int32_t));
    if (n_nodes) *n_nodes = (int)t.nodes.size();
    if (depth) *depth = t.depth;
    return CSG_OK;
}

size_t csg_scene_write(const csg_scene* scene, char* buf, size_t buflen)
{
    if (!scene || scene_within_limits(scene->scene)) return 0;
    const std::string s = write_scene(scene->scene);
    if (buf && buflen) {
        const size_t n = std::min(buflen - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

size_t csg_generate_scene(int n_primitives, uint64_t seed, char* buf, size_t buflen)
{
    const std::string s = generate_scene(n_primitives, seed);
    if (buf && buflen) {
        const size_t n = std::min(buflen - 1, s.size());
        std::memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

float csg_cube_normal_threshold(float half_size, float level) { return cube_normal_threshold_of(half_size, level); }

int csg_scene_set_optimize(csg_scene* scene, int level)
{
    if (!scene) return fail(CSG_ERR_ARG, "null scene");
    scene->scene.optimize = level < 0 ? 0 : level;
    return CSG_OK;
}

// ---- camera / light (host math mirrors Camera.cpp:4-36 and DirectionalLight.h:8-18 operation for operation)
static void cam_normalize(float* v)
{
    float length = (float)std::sqrt((double)(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    if (length > 0) {
        v[0] /= length;
        v[1] /= length;
        v[2] /= length;
    }
}
static void cam_update(csg_camera* c)
{
    const double rx = c->pitch, ry = c->yaw;
    c->forward[0] = (float)(-std::sin(ry) * std::cos(rx));
    c->forward[1] = (float)std::sin(rx);
    c->forward[2] = (float)(-std::cos(ry) * std::cos(rx));
    cam_normalize(c->forward);
    c->right[0] = (float)std::cos(ry);
    c->right[1] = 0;
    c->right[2] = (float)-std::sin(ry);
    cam_normalize(c->right);
    c->up[0] = c->forward[1] * c->right[2] - c->forward[2] * c->right[1];
    c->up[1] = c->forward[2] * c->right[0] - c->forward[0] * c->right[2];
    c->up[2] = c->forward[0] * c->right[1] - c->forward[1] * c->right[0];
    cam_normalize(c->up);
}

void csg_camera_default(csg_camera* cam)
{
    if (!cam) return;
    cam->pos[0] = 0; cam->pos[1] = 0; cam->pos[2] = 5;
    cam->pitch = 0; cam->yaw = 0;
    cam->fov = 90.0f * 3.14159f / 180.0f;
    cam_update(cam);
}

void csg_camera_set(csg_camera* cam, float x, float y, float z, float pitch, float yaw)
{
    if (!cam) return;
    cam->pos[0] = x; cam->pos[1] = y; cam->pos[2] = z;
    cam->pitch = std::fmax(-89.0f * 3.14159f / 180.0f, std::fmin(89.0f * 3.14159f / 180.0f, pitch));
    cam->yaw = yaw;
    cam_update(cam);
}

void csg_camera_set_fov_degrees(csg_camera* cam, float degrees)
{
    if (cam) cam->fov = degrees * 3.14159f / 180.0f;
}

void csg_light_default(csg_light* l)
{
    if (!l) return;
    l->polar = -60.f * 3.14159f / 180.f;
    l->azimuth = -45.f * 3.14159f / 180.f;
}

void csg_light_direction(const csg_light* l, float out3[3])
{
    out3[0] = sinf(l->polar) * cosf(l->azimuth);
    out3[1] = cosf(l->polar);
    out3[2] = sinf(l->polar) * sinf(l->azimuth);
}

// ---- contexts
int csg_upload(const csg_scene* scene, int width, int height, int n_gpus, csg_context** out)
{
    if (n_gpus < 1) return fail(CSG_ERR_ARG, "n_gpus must be >= 1");
    std::vector<int> devs;
    for (int i = 0; i < n_gpus; ++i) devs.push_back(i);
    return create_context(scene, width, height, devs, 0, n_gpus, false, out);
}

int csg_upload_shard(const csg_scene* scene, int width, int height, int device, int shard_rank, int shard_count, csg_context** out)
{
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return fail(CSG_ERR_ARG, "bad shard rank/count");
    std::vector<int> devs{device};
    return create_context(scene, width, height, devs, shard_rank, shard_count, true, out);
}

void csg_free_context(csg_context* c)
{
    if (!c) return;
    if (c->twin) csg_free_context(c->twin);
    if (c->ev_batch0) cudaEventDestroy(c->ev_batch0);
    if (c->ev_batch1) cudaEventDestroy(c->ev_batch1);
    for (Shard& s : c->shards) {
        cudaSetDevice(s.device);
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.copy_stream) {
            cudaStreamSynchronize(s.copy_stream);
            cudaStreamDestroy(s.copy_stream);
            for (cudaEvent_t e : s.ev_band) if (e) cudaEventDestroy(e);
        }
        if (s.ipc_mapped) cudaIpcCloseMemHandle(s.ipc_mapped);
        if (s.owns_local_fb) cudaFree(s.local_fb);
        cudaFree(s.d_exit);
        if (s.h_err) cudaFreeHost(s.h_err);
        cudaFree(s.d_nodes);
        cudaFree(s.d_pool);
        cudaFree(s.d_desc);
        cudaFree(s.d_parent);
        cudaFree(s.d_topo);
        cudaFree(s.d_leaf_boxes);
        cudaFree(s.d_hist);
        cudaFree(s.d_lists);
        cudaFree(s.d_prims);
        cudaFree(s.d_counter);
        if (s.h_tan) cudaFreeHost(s.h_tan);
        if (s.ev_start) cudaEventDestroy(s.ev_start);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (!c->shards.empty()) cudaSetDevice(c->shards[0].device);
    cudaFree(c->d_fb);
    cudaFree(c->d_f32);
    cudaFree(c->d_aov_hit);
    cudaFree(c->d_aov_prim);
    cudaFree(c->d_aov_t);
    cudaFree(c->d_aov_iters);
    cudaGetLastError();
    delete c;
}

int csg_render_enqueue(csg_context* ctx, const csg_camera* cam, const csg_light* light, uint8_t* rgba8_dev)
{
    if (!ctx || !cam || !light) return fail(CSG_ERR_ARG, "null argument");
    float ld[3];
    csg_light_direction(light, ld);
    return enqueue_frame(ctx, cam, ld, OUT_RGBA8, rgba8_dev);
}

int csg_sync(csg_context* ctx)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    return sync_frame(ctx);
}

int csg_last_frame_ms(csg_context* ctx, float* ms)
{
    if (!ctx || !ms) return fail(CSG_ERR_ARG, "null argument");
    if (ctx->frame_pending) {
        int rc = sync_frame(ctx);
        if (rc) return rc;
    }
    *ms = ctx->last_ms;
    return CSG_OK;
}

int csg_stream(csg_context* ctx, void** cuda_stream)
{
    if (!ctx || !cuda_stream || ctx->shards.empty()) return fail(CSG_ERR_ARG, "null argument");
    *cuda_stream = (void*)ctx->shards[0].stream;
    return CSG_OK;
}

uint64_t csg_launch_count(const csg_context* ctx) { return ctx ? ctx->launches + (ctx->twin ? ctx->twin->launches : 0) : 0; }

int csg_framebuffer(csg_context* ctx, uint8_t** rgba8_dev)
{
    if (!ctx || !rgba8_dev) return fail(CSG_ERR_ARG, "null argument");
    *rgba8_dev = ctx->d_fb;
    return CSG_OK;
}

int csg_framebuffer_ipc_handle(csg_context* ctx, void* handle64)
{
    if (!ctx || !handle64) return fail(CSG_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    CU(cudaSetDevice(ctx->shards[0].device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_fb));
    std::memcpy(handle64, &h, 64);
    return CSG_OK;
}

int csg_set_gather_target_ipc(csg_context* ctx, const void* handle64)
{
    if (!ctx || !handle64) return fail(CSG_ERR_ARG, "null argument");
    Shard& s = ctx->shards[0];
    CU(cudaSetDevice(s.device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (s.ipc_mapped) cudaIpcCloseMemHandle(s.ipc_mapped);
    s.ipc_mapped = p;
    s.target = static_cast<uint8_t*>(p);
    // the root's SyncWords sit behind its framebuffer in the same allocation (same width x height on every rank)
    s.sync_words = reinterpret_cast<SyncWords*>(s.target + ctx->sync_off);
    ctx->shard_sync = true;
    return CSG_OK;
}

int csg_set_gather_root(csg_context* ctx, csg_context* root)
{
    if (!ctx || !root) return fail(CSG_ERR_ARG, "null argument");
    if (ctx->width != root->width || ctx->height != root->height || ctx->shard_count != root->shard_count)
        return fail(CSG_ERR_ARG, "the root context must be a shard of the same frame");
    const int root_dev = root->shards[0].device;
    for (Shard& s : ctx->shards) {
        CU(cudaSetDevice(s.device));
        if (s.device != root_dev) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, s.device, root_dev));
            if (!can) return fail(CSG_ERR_CUDA, "device " + std::to_string(s.device) + " cannot access device " + std::to_string(root_dev) + " (P2P)");
            cudaError_t pe = cudaDeviceEnablePeerAccess(root_dev, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(CSG_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
            cudaGetLastError();
        }
        s.target = root->d_fb;
        s.sync_words = reinterpret_cast<SyncWords*>(root->d_fb + root->sync_off);
    }
    ctx->external_target = false;
    ctx->shard_sync = true;
    return CSG_OK;
}

int csg_set_gather_target(csg_context* ctx, uint8_t* rgba8_dev)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    for (Shard& s : ctx->shards) s.target = rgba8_dev ? rgba8_dev : ctx->d_fb;
    ctx->external_target = rgba8_dev != nullptr;
    // one process per GPU: a raw pointer says nothing about where the root keeps its SyncWords — such frames are neither gated
    // nor joined on the device (every rank has to be set up the same way; the caller synchronises the ranks itself)
    if (ctx->multi_process) ctx->shard_sync = rgba8_dev == nullptr;
    return CSG_OK;
}

int csg_pin_host_buffer(void* host, size_t bytes)
{
    if (!host || !bytes) return fail(CSG_ERR_ARG, "null argument");
    CU(cudaHostRegister(host, bytes, cudaHostRegisterPortable));
    return CSG_OK;
}

int csg_unpin_host_buffer(void* host)
{
    if (!host) return fail(CSG_ERR_ARG, "null argument");
    CU(cudaHostUnregister(host));
    return CSG_OK;
}

int csg_read_framebuffer(csg_context* ctx, uint8_t* rgba8_host)
{
    if (!ctx || !rgba8_host) return fail(CSG_ERR_ARG, "null argument");
    int rc = sync_frame(ctx);
    if (rc) return rc;
    CU(cudaMemcpy(rgba8_host, ctx->d_fb, (size_t)ctx->width * ctx->height * 4, cudaMemcpyDeviceToHost));
    return CSG_OK;
}

// Bands of macro-tile rows a host-bound frame is rendered in: the first one must be rendered before any byte moves, the others
// hide behind the copies (PCIe is the bottleneck); every band costs its own pair of launches.  Measured at 4K (68 tile rows) on
// one GPU: 1 band 807 us, 4: 672, 6: 659, 8: 686.  With N shards every shard moves 1/N of the bytes over its own link.
static int host_bands(const csg_context* ctx)
{
    int n = ctx->macro_y >= 16 * ctx->shard_count ? (ctx->macro_y >= 48 ? std::max(1, 6 / ctx->shard_count) : std::max(1, 4 / ctx->shard_count)) : 1;
    if (const char* nb = std::getenv("CSG_B200_BANDS")) n = std::min(std::max(std::atoi(nb), 1), 8);   // tuning aid
    return n;
}

int csg_render(csg_context* ctx, const csg_camera* cam, const csg_light* light, uint8_t* rgba8_out)
{
    if (!ctx || !cam || !light || !rgba8_out) return fail(CSG_ERR_ARG, "null argument");
    const bool dev = is_device_pointer(rgba8_out);
    float ld[3];
    csg_light_direction(light, ld);
    Shard& root = ctx->shards[0];
    const size_t bytes = (size_t)ctx->width * ctx->height * 4;
    if (!dev && !ctx->external_target) {
        // Host output.  The frame is dealt out in macro-tile ROWS (shard mode 1): every shard renders its rows into its own local
        // framebuffer and copies them to the host buffer itself — N GPUs move the frame over N PCIe links (one process per GPU:
        // this rank's rows only; the caller passes every rank the same buffer, e.g. shared memory).  Rows come in bands: band k
        // travels while band k+1 renders.
        const int n_bands = host_bands(ctx);
        const int count = ctx->shard_count;
        const size_t row_bytes = (size_t)kMacroH * ctx->width * 4;
        for (Shard& s : ctx->shards) {
            CU(cudaSetDevice(s.device));
            if (!s.local_fb) {
                CU(cudaMalloc(&s.local_fb, bytes));
                s.owns_local_fb = true;
            }
            if (!s.copy_stream) {
                CU(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
                for (cudaEvent_t& e : s.ev_band) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
        }
        int rc = CSG_OK;
        int rm[4] = {0, 0, 0, 0};   // union of the bands' traced rectangles, for csg_prune_stats
        for (int b = 0; b < n_bands && !rc; ++b) {
            ctx->band_m0 = ctx->macro_y * b / n_bands;
            ctx->band_m1 = ctx->macro_y * (b + 1) / n_bands;
            rc = enqueue_frame(ctx, cam, ld, OUT_RGBA8, nullptr, 1);
            if (rc) break;
            if (ctx->last_rm[2] > 0 && ctx->last_rm[3] > 0) {
                if (rm[3] == 0) { rm[0] = ctx->last_rm[0]; rm[1] = ctx->last_rm[1]; rm[2] = ctx->last_rm[2]; }
                rm[3] = ctx->last_rm[1] + ctx->last_rm[3] - rm[1];
            }
            for (Shard& s : ctx->shards) {
                // this shard's rows of the band: first, first + count, ...; all but a ragged last row of the frame in one 2-D copy
                const int first = shard_row_first(ctx->band_m0, s.rank, count);
                int n_rows = shard_row_count(ctx->band_m0, ctx->band_m1 - ctx->band_m0, s.rank, count);
                if (n_rows == 0) continue;
                CU(cudaSetDevice(s.device));
                CU(cudaEventRecord(s.ev_band[b], s.stream));
                CU(cudaStreamWaitEvent(s.copy_stream, s.ev_band[b], 0));
                const int last = first + (n_rows - 1) * count;
                const size_t last_lines = std::min<size_t>((size_t)(last + 1) * kMacroH, (size_t)ctx->height) - (size_t)last * kMacroH;
                if (last_lines < (size_t)kMacroH) {
                    const size_t off = (size_t)last * row_bytes;
                    CU(cudaMemcpyAsync(rgba8_out + off, s.local_fb + off, last_lines * ctx->width * 4, cudaMemcpyDeviceToHost, s.copy_stream));
                    --n_rows;
                }
                if (n_rows > 0) {
                    const size_t off = (size_t)first * row_bytes;
                    static const bool copy2d = !(std::getenv("CSG_B200_COPY2D") && std::getenv("CSG_B200_COPY2D")[0] == '0');   // tuning aid
                    if (count == 1) CU(cudaMemcpyAsync(rgba8_out + off, s.local_fb + off, n_rows * row_bytes, cudaMemcpyDeviceToHost, s.copy_stream));
                    else if (copy2d) CU(cudaMemcpy2DAsync(rgba8_out + off, count * row_bytes, s.local_fb + off, count * row_bytes, row_bytes, (size_t)n_rows,
                                                          cudaMemcpyDeviceToHost, s.copy_stream));
                    else
                        for (int k = 0; k < n_rows; ++k)
                            CU(cudaMemcpyAsync(rgba8_out + off + (size_t)k * count * row_bytes, s.local_fb + off + (size_t)k * count * row_bytes, row_bytes,
                                               cudaMemcpyDeviceToHost, s.copy_stream));
                }
            }
        }
        ctx->band_m0 = ctx->band_m1 = 0;
        for (int i = 0; i < 4; ++i) ctx->last_rm[i] = rm[i];
        if (rc) return rc;
        for (Shard& s : ctx->shards) {
            CU(cudaSetDevice(s.device));
            CU(cudaStreamSynchronize(s.copy_stream));
        }
        return sync_frame(ctx);
    }
    // device destination (or an external gather target): multi-shard contexts gather into the root framebuffer first
    const bool direct = dev && ctx->shards.size() == 1 && ctx->shard_count == 1;
    int rc = enqueue_frame(ctx, cam, ld, OUT_RGBA8, direct ? (void*)rgba8_out : nullptr);
    if (rc) return rc;
    if (!direct) {
        CU(cudaSetDevice(root.device));
        CU(cudaMemcpyAsync(rgba8_out, root.target, bytes, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, root.stream));
        CU(cudaStreamSynchronize(root.stream));
    }
    return sync_frame(ctx);
}

int csg_render_batch(csg_context* ctx, const csg_camera* cams, int n_frames, const csg_light* light, uint8_t* rgba8_out)
{
    if (!ctx || !cams || !light || !rgba8_out || n_frames < 1) return fail(CSG_ERR_ARG, "bad argument");
    if (ctx->shards.size() != 1 || ctx->multi_process) return fail(CSG_ERR_ARG, "csg_render_batch needs a single-GPU context");
    Shard& s0 = ctx->shards[0];
    CU(cudaSetDevice(s0.device));
    if (!ctx->twin) {
        std::vector<int> devs{s0.device};
        int rc = create_context(&ctx->scene_copy, ctx->width, ctx->height, devs, 0, 1, false, &ctx->twin);
        if (rc) return rc;
        CU(cudaEventCreate(&ctx->ev_batch0));
        CU(cudaEventCreate(&ctx->ev_batch1));
    }
    csg_context* slot[2] = {ctx, ctx->twin};
    ctx->twin->ss = ctx->ss;
    ctx->twin->prune = ctx->prune;
    ctx->twin->prune_flat = ctx->prune_flat;
    const bool dev = is_device_pointer(rgba8_out);
    const size_t bytes = (size_t)ctx->width * ctx->height * 4;
    float ld[3];
    csg_light_direction(light, ld);
    // the device tanf(fov/2) is fetched with a blocking round trip the first time a field of view is seen: do that up front
    for (int k = 0; k < n_frames; ++k)
        for (csg_context* c : slot)
            if (!(cams[k].fov == c->cached_fov)) {
                float t;
                int rc = csg_device_tan_half_fov(c, cams[k].fov, &t);
                if (rc) return rc;
                c->cached_fov = cams[k].fov; c->cached_tan = t;
            }
    CU(cudaEventRecord(ctx->ev_batch0, s0.stream));
    CU(cudaStreamWaitEvent(ctx->twin->shards[0].stream, ctx->ev_batch0, 0));
    for (int k = 0; k < n_frames; ++k) {
        csg_context* c = slot[k & 1];
        uint8_t* dst = rgba8_out + (size_t)k * bytes;
        int rc = enqueue_frame(c, &cams[k], ld, OUT_RGBA8, dev ? (void*)dst : nullptr);
        if (rc) return rc;
        if (!dev) CU(cudaMemcpyAsync(dst, c->d_fb, bytes, cudaMemcpyDeviceToHost, c->shards[0].stream));
    }
    cudaStream_t t1 = ctx->twin->shards[0].stream;
    CU(cudaEventRecord(ctx->ev_batch1, t1));
    CU(cudaStreamWaitEvent(s0.stream, ctx->ev_batch1, 0));
    CU(cudaEventRecord(ctx->ev_batch1, s0.stream));
    CU(cudaEventSynchronize(ctx->ev_batch1));
    CU(cudaStreamSynchronize(t1));
    CU(cudaEventElapsedTime(&ctx->last_ms, ctx->ev_batch0, ctx->ev_batch1));
    ctx->frame_pending = false;
    ctx->twin->frame_pending = false;
    CU(cudaGetLastError());
    return CSG_OK;
}

int csg_render_f32(csg_context* ctx, const csg_camera* cam, const csg_light* light, float* rgba_f32_out)
{
    if (!ctx || !cam || !light || !rgba_f32_out) return fail(CSG_ERR_ARG, "null argument");
    if (ctx->shards.size() != 1) return fail(CSG_ERR_ARG, "csg_render_f32 needs a single-GPU context");
    Shard& root = ctx->shards[0];
    CU(cudaSetDevice(root.device));
    const bool dev = is_device_pointer(rgba_f32_out);
    const size_t bytes = (size_t)ctx->width * ctx->height * 16;
    float ld[3];
    csg_light_direction(light, ld);
    if (!dev && !ctx->d_f32) CU(cudaMalloc(&ctx->d_f32, bytes));
    int rc = enqueue_frame(ctx, cam, ld, OUT_F32, dev ? rgba_f32_out : ctx->d_f32);
    if (rc) return rc;
    if (!dev) {
        CU(cudaMemcpyAsync(rgba_f32_out, ctx->d_f32, bytes, cudaMemcpyDeviceToHost, root.stream));
        CU(cudaStreamSynchronize(root.stream));
    }
    return sync_frame(ctx);
}

int csg_render_stats(csg_context* ctx, const csg_camera* cam, int32_t* iterations)
{
    if (!ctx || !cam || !iterations) return fail(CSG_ERR_ARG, "null argument");
    if (ctx->shards.size() != 1) return fail(CSG_ERR_ARG, "csg_render_stats needs a single-GPU context");
    CU(cudaSetDevice(ctx->shards[0].device));
    const size_t n = (size_t)ctx->width * ctx->height;
    if (!ctx->d_aov_iters) CU(cudaMalloc(&ctx->d_aov_iters, n * 4));
    int rc = csg_render_aov(ctx, cam, nullptr, nullptr, nullptr);
    if (rc) return rc;
    CU(cudaMemcpy(iterations, ctx->d_aov_iters, n * 4, cudaMemcpyDeviceToHost));
    return CSG_OK;
}

int csg_render_aov(csg_context* ctx, const csg_camera* cam, uint8_t* hit, int32_t* prim_id, float* t)
{
    if (!ctx || !cam) return fail(CSG_ERR_ARG, "null argument");
    if (ctx->shards.size() != 1) return fail(CSG_ERR_ARG, "csg_render_aov needs a single-GPU context");
    Shard& root = ctx->shards[0];
    CU(cudaSetDevice(root.device));
    const size_t n = (size_t)ctx->width * ctx->height;
    if (!ctx->d_aov_hit) {
        CU(cudaMalloc(&ctx->d_aov_hit, n));
        CU(cudaMalloc(&ctx->d_aov_prim, n * 4));
        CU(cudaMalloc(&ctx->d_aov_t, n * 4));
    }
    int rc = enqueue_frame(ctx, cam, nullptr, OUT_AOV, nullptr);
    if (rc) return rc;
    rc = sync_frame(ctx);
    if (rc) return rc;
    if (hit) CU(cudaMemcpy(hit, ctx->d_aov_hit, n, cudaMemcpyDeviceToHost));
    if (prim_id) CU(cudaMemcpy(prim_id, ctx->d_aov_prim, n * 4, cudaMemcpyDeviceToHost));
    if (t) CU(cudaMemcpy(t, ctx->d_aov_t, n * 4, cudaMemcpyDeviceToHost));
    return CSG_OK;
}

int csg_set_view_cache(csg_context* ctx, int enabled)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    ctx->view_cache = enabled != 0;
    for (Shard& s : ctx->shards) s.last_q_valid = false;
    return CSG_OK;
}

int csg_set_pruning(csg_context* ctx, int enabled)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    ctx->prune = enabled != 0 && ctx->prune_alloc;
    ctx->prune_flat = enabled != 2;   // 2: the tree-walking csg_prune_kernel
    for (Shard& s : ctx->shards) s.last_q_valid = false;
    return CSG_OK;
}

int csg_prune_stats(csg_context* ctx, int* traced_tiles, int* empty_tiles, int* fallback_tiles, long long* pruned_nodes)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    int rc = sync_frame(ctx);
    if (rc) return rc;
    Shard& s = ctx->shards[0];
    CU(cudaSetDevice(s.device));
    std::vector<TileDesc> d((size_t)std::max(s.n_slots, 1));
    CU(cudaMemcpy(d.data(), s.d_desc, d.size() * sizeof(TileDesc), cudaMemcpyDeviceToHost));
    int traced = 0, empty = 0, fb = 0;
    long long nodes = 0;
    // the traced rectangle of the last frame, in this shard's slots
    for (int jy = 0; jy < ctx->last_rm[3]; ++jy)
        for (int jx = 0; jx < ctx->last_rm[2]; ++jx) {
            const int j = jy * ctx->last_rm[2] + jx;
            const int mx = ctx->last_rm[0] + jx, my = ctx->last_rm[1] + jy;
            if ((ctx->last_mode ? my : j) % ctx->shard_count != s.rank) continue;
            const TileDesc& t = d[(size_t)slot_of_macro(ctx->last_mode, mx, my, ctx->macro_x, ctx->shard_count)];
            ++traced;
            if (t.n_nodes == 0) ++empty;
            else if (t.offset32 == 0 && ctx->prune) ++fb;
            else nodes += t.n_nodes;
        }
    if (traced_tiles) *traced_tiles = traced;
    if (empty_tiles) *empty_tiles = empty;
    if (fallback_tiles) *fallback_tiles = fb;
    if (pruned_nodes) *pruned_nodes = nodes;
    return CSG_OK;
}

int csg_set_supersampling(csg_context* ctx, int samples_per_axis)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    if (samples_per_axis < 1 || samples_per_axis > 16) return fail(CSG_ERR_ARG, "samples_per_axis must be in [1, 16]");
    if ((long long)ctx->width * samples_per_axis > 16777216ll || (long long)ctx->height * samples_per_axis > 16777216ll)
        return fail(CSG_ERR_ARG, "virtual grid too large for exact float pixel coordinates");
    ctx->ss = samples_per_axis;
    return CSG_OK;
}

int csg_device_tan_half_fov(csg_context* ctx, float fov, float* out)
{
    if (!ctx || !out) return fail(CSG_ERR_ARG, "null argument");
    int rc = device_tan(ctx, fov, out);
    if (rc) return rc;
    return CSG_OK;
}

int csg_fp32_peak_tflops(int device, float* tflops)
{
    if (!tflops) return fail(CSG_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(CSG_ERR_NO_DEVICE, "no such CUDA device");
    }
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float* d = nullptr;
    CU(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 0.f;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        csg_ffma_probe_kernel<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, (float)(flop / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return CSG_OK;
}

#ifdef CSG_FRAME_PROBE
int csg_debug_sync_probe(int device, void* out8)
{
    CU(cudaSetDevice(device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out8, g_sync_probe, sizeof(unsigned long long) * 8));
    return CSG_OK;
}
int csg_debug_frame_probe(void* out, size_t bytes)
{
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, g_frame_probe, std::min(bytes, sizeof(unsigned long long) * 8192 * 8)));
    void* sym = nullptr;
    CU(cudaGetSymbolAddress(&sym, g_frame_probe));
    CU(cudaMemset(sym, 0, sizeof(unsigned long long) * 8192 * 8));
    return CSG_OK;
}
#endif

#ifdef CSG_PRUNE_PROBE
int csg_debug_prune_probe(void* out, size_t bytes)
{
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, g_prune_probe, std::min(bytes, sizeof(unsigned long long) * 4096 * 16)));
    void* sym = nullptr;
    CU(cudaGetSymbolAddress(&sym, g_prune_probe));
    CU(cudaMemset(sym, 0, sizeof(unsigned long long) * 4096 * 16));
    return CSG_OK;
}
#endif

const char* csg_context_info(csg_context* ctx) { return ctx ? ctx->info.c_str() : ""; }

int csg_shard_tile(int macro_x, int macro_y, int rm_x0, int rm_y0, int rm_w, int rm_h, int shard_mode, int shard_rank, int shard_count,
                   int tile, int* mx, int* my, int* slot, int* n_tiles, int* n_slots)
{
    if (macro_x < 1 || macro_y < 1 || rm_w < 0 || rm_h < 0 || rm_x0 < 0 || rm_y0 < 0 || rm_x0 + rm_w > macro_x || rm_y0 + rm_h > macro_y ||
        shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count)
        return fail(CSG_ERR_ARG, "bad tile rectangle or shard");
    const long long traced = (long long)rm_w * rm_h;
    const long long mine = shard_mode ? (long long)shard_row_count(rm_y0, rm_h, shard_rank, shard_count) * rm_w
                                      : (traced > shard_rank ? (traced - shard_rank + shard_count - 1) / shard_count : 0);
    if (n_tiles) *n_tiles = (int)mine;
    if (n_slots) *n_slots = slots_per_shard(macro_x, macro_y, shard_count);
    if (tile < 0 || tile >= mine) return fail(CSG_ERR_ARG, "tile out of range");
    // the same host-computed reciprocal the kernels divide with (fill_params)
    const unsigned int magic = (rm_w > 1 && rm_w < 4096 && traced < (1ll << 20)) ? (unsigned int)((1ull << 32) / (unsigned long long)rm_w + 1ull) : 0u;
    int x, y;
    shard_tile_coords(tile, shard_mode, shard_rank, shard_count, rm_x0, rm_y0, rm_w, magic, shard_row_first(rm_y0, shard_rank, shard_count), x, y);
    if (mx) *mx = x;
    if (my) *my = y;
    if (slot) *slot = slot_of_macro(shard_mode, x, y, macro_x, shard_count);
    return CSG_OK;
}

}  // extern "C"
