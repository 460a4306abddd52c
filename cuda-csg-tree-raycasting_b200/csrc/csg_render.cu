// csg_render.cu — frame kernel, launcher and the C ABI of libcsg_b200 (include/csg_b200.h).
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
#include <string>
#include <vector>

#include "../../include/csg_b200.h"
#include "csg_kernel.cuh"
#include "csg_scene.h"

using namespace csgb;

// =========================================================================================== device code
namespace csgb {

// CTA shapes the frame kernel is compiled for.  All three keep 24 warps per SM at <= 80 registers/thread
// (768 threads x 80 registers x 1 CTA, 384 x 2, 256 x 3 all fill the 64K-register file); they differ in how many copies
// of the staged tree and how much stack an SM holds.  The launcher picks the shape with the most resident warps for the
// scene at hand, the largest CTA on ties (one tree copy per SM).
constexpr int kShapes = 3;
constexpr int kShapeThreads[kShapes] = {768, 384, 256};
constexpr int min_blocks_for(int threads) { return threads >= 768 ? 1 : threads >= 384 ? 2 : 3; }
constexpr int kWarpTileW = 8, kWarpTileH = 4;   // one warp = one 8x4 pixel tile (Raycaster.cuh:7-8 uses the same shape)
constexpr int kMacroW = 64, kMacroH = 32;       // sharding unit: 8x8 warp tiles

// Hit details + Phong (sphere/cylinder/cubeHitDetails :183-200/:338-372/:436-457 and LightningKernel :49-111).
__device__ __noinline__ float4 shade_pixel(const Hit res, const Ray r, const float4* __restrict__ prims, const FrameParams& p, const float* __restrict__ s_light)
{
    if (is_miss(res)) return make_float4(0.08f, 0.08f, 0.11f, 1.0f);   // :109
    const uint32_t id = (res.m & H_META_MASK) >> H_ID_SHIFT;
    const uint32_t kind = (res.m >> H_KIND_SHIFT) & 7u;
    const float4 col = __ldg(&prims[id * 5 + 0]);
    const float4 pc = __ldg(&prims[id * 5 + 1]);
    const float t = res.t;
    const float px = __fmaf_rn(t, r.dx, r.ox), py = __fmaf_rn(t, r.dy, r.oy), pz = __fmaf_rn(t, r.dz, r.oz);
    float nx, ny, nz;
    if (kind == 3u) {                                   // sphereHitDetails :189-195
        nx = px - pc.x; ny = py - pc.y; nz = pz - pc.z;
    } else if (kind == 5u) {                            // cubeHitDetails :446-451
        const float bias = 1.00001f;
        nx = (float)__float2int_rz(__fmul_rn(__fdiv_rn(px - pc.x, pc.w), bias));
        ny = (float)__float2int_rz(__fmul_rn(__fdiv_rn(py - pc.y, pc.w), bias));
        nz = (float)__float2int_rz(__fmul_rn(__fdiv_rn(pz - pc.z, pc.w), bias));
    } else {                                            // cylinderHitDetails :345-365
        const float4 pb = __ldg(&prims[id * 5 + 2]);
        const float4 pv = __ldg(&prims[id * 5 + 3]);
        if (res.m & H_FLAG1) { nx = -pv.x; ny = -pv.y; nz = -pv.z; }
        else if (res.m & H_FLAG2) { nx = pv.x; ny = pv.y; nz = pv.z; }
        else {
            const float ocx = r.ox - pb.x, ocy = r.oy - pb.y, ocz = r.oz - pb.z;
            const float dV = dot_ref(pv.x, pv.y, pv.z, r.dx, r.dy, r.dz);
            const float ocv = dot_ref(pv.x, pv.y, pv.z, ocx, ocy, ocz);
            const float m = __fmaf_rn(t, dV, ocv);
            nx = __fmaf_rn(-pv.x, m, px - pb.x);
            ny = __fmaf_rn(-pv.y, m, py - pb.y);
            nz = __fmaf_rn(-pv.z, m, pz - pb.z);
        }
    }
    if (!((res.m & (H_FLAG1 | H_FLAG2)) && kind == 4u)) {   // caps carry the unit axis as is; everything else is normalised
        const float inv = __frcp_rn(__fsqrt_rn(dot_ref(nx, ny, nz, nx, ny, nz)));
        nx *= inv; ny *= inv; nz *= inv;
    }
    if (res.m & H_FLIP) { nx = -nx; ny = -ny; nz = -nz; }
    if ((res.m & H_CLS) == H_EXIT) { nx = -nx; ny = -ny; nz = -nz; }

    // LightningKernel :78-103
    const float Lx = s_light[0], Ly = s_light[1], Lz = s_light[2];   // normalize(lightDir), once per CTA
    float vx = p.cam_pos[0] - px, vy = p.cam_pos[1] - py, vz = p.cam_pos[2] - pz;
    const float iv = __frcp_rn(__fsqrt_rn(dot_ref(vx, vy, vz, vx, vy, vz)));
    vx *= iv; vy *= iv; vz *= iv;
    const float in = __frcp_rn(__fsqrt_rn(dot_ref(nx, ny, nz, nx, ny, nz)));   // reflect() re-normalises n
    const float ux = nx * in, uy = ny * in, uz = nz * in;
    const float dn = __fmaf_rn(-Lz, uz, __fmaf_rn(-Lx, ux, uy * -Ly));
    const float two = dn + dn;
    const float rx = __fmaf_rn(-ux, two, -Lx), ry = __fmaf_rn(-uy, two, -Ly), rz = __fmaf_rn(-uz, two, -Lz);
    const float diff = fmaxf(dot_ref(nx, ny, nz, Lx, Ly, Lz), 0.0f);
    const float sb = fmaxf(dot_ref(vx, vy, vz, rx, ry, rz), 0.0f);
    const float spec = powf(sb, 30.0f);
    const float k = __fmaf_rn(spec, 0.7f, __fmaf_rn(diff, 0.8f, 0.2f));
    float4 o;
    o.x = fminf(fmaxf(col.x * k, 0.0f), 1.0f);
    o.y = fminf(fmaxf(col.y * k, 0.0f), 1.0f);
    o.z = fminf(fmaxf(col.z * k, 0.0f), 1.0f);
    o.w = 1.0f;
    return o;
}

__device__ __forceinline__ uint32_t to_u8(float c)
{  // Q12: (int)(clamp(c,0,1)*255 + 0.5)
    return (uint32_t)__float2int_rz(__fadd_rn(__fmul_rn(fminf(fmaxf(c, 0.0f), 1.0f), 255.0f), 0.5f));
}

// CSGRayCast (RaycastingKernels.cu:459-512) re-expressed as an explicit-frame evaluation; equivalence with the
// reference's GoTo/Compute/SaveLft action machine is argued in DESIGN.md §"State machine".
//
// One 16-byte frame per operator on the current path, in shared memory ([level][thread]):
//   x = saved tmin (F_FIRST_*) or saved hit t (F_LOAD_*),  y = saved hit meta | return state,
//   z = F_FIRST_*: lower bound of the pending sibling's hits (prune test),  w = byte offset of the operator's record.
//
// Additions over the reference's traversal order, all result-preserving (DESIGN.md §"Culling contract"):
//   * a Union evaluates the child whose box the ray enters first; Difference/Intersection keep left-first;
//   * when the first child returns a hit at t and the pending sibling's box starts beyond t, the sibling cannot change
//     the outcome (Union: every cell with a farther Enter or a Miss on the other side returns this hit; Difference: same
//     for the right operand) and is skipped;
//   * nearest-Enter search (ST_SEARCH): a "pure" subtree (Unions over spheres/cubes only) whose box lies ahead of tmin is
//     evaluated as a closest-hit BVH search with a shrinking limit instead of the frame machine.  With every leaf result an
//     Enter or a Miss, each Union of the subtree returns the nearer Enter (EE lt/gt, EM, ME cells), i.e. the subtree returns
//     its globally nearest Enter; the search aborts — and the subtree is re-evaluated by the frame machine — as soon as a
//     leaf reports an Exit or two leaves tie for the nearest hit (the only inputs on which the cells differ from "min").
//     The limit starts at the hit already known on the other side of the parent when every farther Enter (or a Miss) of
//     this side gives the same parent outcome (Union either side, Difference right side): such results are equivalent, so
//     subtrees beyond the limit need not be looked at.
constexpr uint32_t kSearchMark = 0xfffffffeu;

template <bool COUNT>
__device__ __forceinline__ Hit traverse(const unsigned char* __restrict__ tree, const float4* __restrict__ prims,
                                        const uint32_t* __restrict__ table, const uint32_t stack,
                                        const uint32_t stack_stride, const Ray& r, const bool root_is_leaf, const bool root_pure, const bool root_gated, int& iters)
{
    enum { ST_ENTER = 0, ST_SEARCH = 1, ST_LOOPL = 2, ST_LOOPR = 3, ST_COMPUTE = 4, ST_RETURN = 5, ST_DONE = 6 };
    Hit L = make_miss(), R = make_miss();      // ST_SEARCH: L = nearest Enter so far, R.t = limit
    float tmin = 0.0f;                        // :466
    if (root_is_leaf) {
        // The scene's root is a primitive: GoTo's leaf branch on the virtual root, no box test (:582-594, Q7).
        // A pruned tile tree that collapsed to one primitive (root_gated): the primitive is still reached through its
        // operators in the reference, so a cylinder keeps its gating box (Q6).
        bool go;
        float tn;
        uint32_t cm;
        eval_child(tree, prims, 0u, r, tmin, root_gated, L, go, tn, cm);
        return L;
    }
    uint32_t n = 0u;                           // byte offset of the current operator's record
    uint32_t sp = stack;                       // next free frame (shared-memory address; stack_stride = bytes between levels)
    sts128(sp, make_uint4(0u, 0u, 0u, 0xffffffffu)); // sentinel frame: popping it ends the traversal (no base pointer to keep)
    sp += stack_stride;
    int st = ST_ENTER;
    if (root_pure) {                           // the whole scene is one pure subtree
        sts128(sp, make_uint4(0u, 0u, 0u, kSearchMark));
        sp += stack_stride;
        R.t = INFINITY;
        st = ST_SEARCH;
    }
    while (st != ST_DONE) {
        if (COUNT) iters += (st == ST_SEARCH) ? (1 << 20) : (st == ST_ENTER) ? (1 << 10) : 1;   // packed: search visits | frame-machine visits | other iterations
        if (st <= ST_LOOPR) {
            const uint32_t meta = *reinterpret_cast<const uint32_t*>(tree + n + 28);
            const uint32_t op = meta & 7u;
            const uint32_t cl = n + 32u, cr = (meta >> 8) << 5;
            Hit a = make_miss(), b = make_miss();
            bool goA = false, goB = false;
            float tnA = -INFINITY, tnB = -INFINITY;
            uint32_t mA = 0u, mB = 0u;
            if (st != ST_LOOPR) eval_child(tree, prims, cl, r, tmin, st <= ST_SEARCH, a, goA, tnA, mA);
            if (st == ST_ENTER && op != 0u && !goA && is_miss(a)) {
                // left operand of a Difference/Intersection already missed: the node's result is Miss whatever the right
                // operand does (all M* cells of both tables, :670-677) — skip the right subtree (Q8)
                L = a; R = a;
                st = ST_RETURN;
            } else {
                if (st != ST_LOOPL) eval_child(tree, prims, cr, r, tmin, st <= ST_SEARCH, b, goB, tnB, mB);
                if (st == ST_LOOPL) { L = a; st = ST_COMPUTE; }
                else if (st == ST_LOOPR) { R = b; st = ST_COMPUTE; }
                else if (st == ST_SEARCH) {
                    // leaf results are candidates; an Exit or a tie for the nearest hit ends the search
                    bool abort = false;
                    float lim = R.t;
                    if (!is_miss(a)) {
                        if ((a.m & H_CLS) == H_EXIT) abort = true;
                        else if (a.t < lim) { L = a; lim = a.t; }
                        else if (a.t == lim) { if (is_miss(L)) L = a; else abort = true; }
                    }
                    if (!is_miss(b)) {
                        if ((b.m & H_CLS) == H_EXIT) abort = true;
                        else if (b.t < lim) { L = b; lim = b.t; }
                        else if (b.t == lim) { if (is_miss(L)) L = b; else abort = true; }
                    }
                    R.t = lim;
                    if (abort) {                 // back to the subtree's root, this time through the frame machine
                        uint4 f;
                        do { sp -= stack_stride; f = lds128(sp); } while (f.w != kSearchMark);
                        n = f.z; st = ST_ENTER;
                    } else {
                        goA = goA && !(tnA > lim);
                        goB = goB && !(tnB > lim);
                        if (goA && goB) {
                            const bool right_first = tnB < tnA;
                            sts128(sp, make_uint4(__float_as_uint(right_first ? tnA : tnB), 0u, 0u, right_first ? cl : cr));
                            sp += stack_stride; n = right_first ? cr : cl;
                        } else if (goA) { n = cl; }
                        else if (goB) { n = cr; }
                        else {
                            for (;;) {
                                sp -= stack_stride;
                                const uint4 f = lds128(sp);
                                if (f.w == kSearchMark) { R = L; st = ST_RETURN; break; }   // the subtree's result: nearest Enter or Miss
                                if (!(__uint_as_float(f.x) > lim)) { n = f.w; break; }
                            }
                        }
                    }
                } else {
                    L = a; R = b;
                    // sibling pruning against a leaf hit that is already known
                    if (op != 2u) {
                        if (goB && !goA && !is_miss(L) && tnB > L.t) goB = false;
                        if (op == 0u && goA && !goB && !is_miss(R) && tnA > R.t) goA = false;
                    }
                    if (!goA && !goB) {
                        st = ST_COMPUTE;                                                   // :578
                    } else {
                        uint32_t first, fm;   // subtree to descend into now, and its meta word
                        float ftn, lim = INFINITY;
                        if (!goA) {                                                        // :556-561
                            sts128(sp, make_uint4(__float_as_uint(L.t), L.m | F_LOAD_LFT, 0u, n));
                            first = cr; fm = mB; ftn = tnB;
                            if (op != 2u && !is_miss(L)) lim = L.t;
                        } else if (!goB) {                                                 // :562-567
                            sts128(sp, make_uint4(__float_as_uint(R.t), R.m | F_LOAD_RGH, 0u, n));
                            first = cl; fm = mA; ftn = tnA;
                            if (op == 0u && !is_miss(R)) lim = R.t;
                        } else {                                                           // :568-574
                            const bool right_first = (op == 0u) && (tnB < tnA);
                            const uint32_t pend_pure = ((right_first ? mA : mB) >> 6) & 1u;
                            sts128(sp, make_uint4(__float_as_uint(tmin), (right_first ? F_FIRST_RGH : F_FIRST_LFT) | pend_pure,
                                             __float_as_uint(right_first ? tnA : tnB), n));
                            first = right_first ? cr : cl; fm = right_first ? mB : mA; ftn = right_first ? tnB : tnA;
                        }
                        sp += stack_stride; n = first;
                        if ((fm & kMetaPure) && ftn > tmin) {   // pure subtree ahead of tmin: nearest-Enter search
                            sts128(sp, make_uint4(0u, 0u, first, kSearchMark));
                            sp += stack_stride;
                            L = make_miss(); R.t = lim; st = ST_SEARCH;
                        }
                    }
                }
            }
        }
        if (st == ST_COMPUTE) {                                                        // Compute :597-661
            const uint32_t meta = *reinterpret_cast<const uint32_t*>(tree + n + 28);
            const uint32_t op = meta & 7u;
            const uint32_t e = table[op * 9u + (L.m & H_CLS) * 3u + (R.m & H_CLS)];
            const uint32_t o = (L.t < R.t) ? (e & 7u) : (L.t > R.t) ? ((e >> 3) & 7u) : ((e >> 6) & 7u);
            if (o == O_RETL) { R = L; st = ST_RETURN; }
            else if (o == O_RETR || o == O_RETR_FLIP) {
                if (o == O_RETR_FLIP) R.m ^= (H_FLIP | 1u);                            // :629-635 toggles Flip and Enter<->Exit
                L = R; st = ST_RETURN;
            } else if (o == O_LOOPL) {                                                 // :640-646
                tmin = L.t;
                if (meta & kMetaLeftLeaf) st = ST_LOOPL;
                else { sts128(sp, make_uint4(__float_as_uint(R.t), R.m | F_LOAD_RGH, 0u, n)); sp += stack_stride; n = n + 32u; st = ST_ENTER; }
            } else if (o == O_LOOPR) {                                                 // :647-653
                tmin = R.t;
                if (meta & kMetaRightLeaf) st = ST_LOOPR;
                else { sts128(sp, make_uint4(__float_as_uint(L.t), L.m | F_LOAD_LFT, 0u, n)); sp += stack_stride; n = (meta >> 8) << 5; st = ST_ENTER; }
            } else { L = R = make_miss(); st = ST_RETURN; }                            // :654-660
        }
        if (st == ST_RETURN) {                  // action = actionStack.pop(); node = GetParent() (:624-625 etc.); L == R == result
            sp -= stack_stride;
            const uint4 f = lds128(sp);
            if (f.w == 0xffffffffu) { st = ST_DONE; }
            else {
                n = f.w;
                const uint32_t ret = f.y & F_RET_MASK;
                if (ret == F_LOAD_LFT) {        // :611-614
                    L.t = __uint_as_float(f.x); L.m = f.y & H_META_MASK; st = ST_COMPUTE;
                } else if (ret == F_LOAD_RGH) { // :615-618
                    R.t = __uint_as_float(f.x); R.m = f.y & H_META_MASK; st = ST_COMPUTE;
                } else {                        // SaveLft :476-481: restore tmin, keep the first result, evaluate the sibling
                    tmin = __uint_as_float(f.x);
                    const uint32_t pm = *reinterpret_cast<const uint32_t*>(tree + n + 28);
                    const uint32_t pop = pm & 7u;
                    const bool miss = is_miss(L);
                    const float ptn = __uint_as_float(f.z);     // entry distance of the pending sibling's box (-inf: not a bound)
                    if (miss ? (pop != 0u) : (pop != 2u && ptn > L.t)) {
                        // Difference/Intersection whose left operand missed -> Miss; or the sibling lies beyond this hit -> this hit.
                        // Either way the node's result is what L == R already hold; stay in ST_RETURN.
                    } else {
                        uint32_t sib;
                        if (ret == F_FIRST_LFT) {
                            sts128(sp, make_uint4(__float_as_uint(L.t), L.m | F_LOAD_LFT, 0u, n));
                            sib = (pm >> 8) << 5;
                        } else {
                            sts128(sp, make_uint4(__float_as_uint(R.t), R.m | F_LOAD_RGH, 0u, n));
                            sib = n + 32u;
                        }
                        sp += stack_stride; n = sib; st = ST_ENTER;
                        if ((f.y & 1u) && ptn > tmin) {
                            const float lim = (pop != 2u && !miss) ? L.t : INFINITY;
                            sts128(sp, make_uint4(0u, 0u, sib, kSearchMark));
                            sp += stack_stride;
                            L = make_miss(); R.t = lim; st = ST_SEARCH;
                        }
                    }
                }
            }
        }
    }
    return L;                                   // :511
}

template <int MODE, int kThreads, bool kSuper>   // kSuper: more than one ray per pixel (keeps the sample loop and its accumulators out of the common case)
__global__ void __launch_bounds__(kThreads, min_blocks_for(kThreads)) csg_frame_kernel(const __grid_constant__ FrameParams p)
{
    // shared memory: [outcome table 128 B][stack: (levels+2) x kThreads x 16 B][per-warp tree copy: warp_tree_nodes x 32 B].
    // Every 64x32-pixel macro tile has its own pruned, origin-relative tree (csg_prune_kernel), a few hundred bytes to a few
    // KB; a warp copies the tree of its current tile into shared memory when it fits, and reads it through L1 otherwise.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_table = reinterpret_cast<uint32_t*>(smem_raw);
    uint4* s_stack = reinterpret_cast<uint4*>(smem_raw + 128);

    const int tid = threadIdx.x, lane = tid & 31;
    const float ox = p.cam_pos[0], oy = p.cam_pos[1], oz = p.cam_pos[2];

    if (tid < 27) s_table[tid] = kOutcomeTable[tid];
    if (tid == 27) {   // L = normalize(lightDir), LightningKernel :78: the same for every pixel of the frame
        const float il = __frcp_rn(__fsqrt_rn(dot_ref(p.light[0], p.light[1], p.light[2], p.light[0], p.light[1], p.light[2])));
        float* s_light = reinterpret_cast<float*>(s_table + 28);
        s_light[0] = p.light[0] * il; s_light[1] = p.light[1] * il; s_light[2] = p.light[2] * il;
    }
    __syncthreads();
    const uint32_t my_stack = (uint32_t)__cvta_generic_to_shared(s_stack + tid);   // frames are addressed in the shared window: 32-bit
    // per-warp copy of the current tile's tree (when it fits): traversal then reads shared memory instead of L1/L2
    const uint32_t my_tree_off = 128u + 16u * (uint32_t)((p.stack_levels + 2) * kThreads + (tid >> 5) * (2 * p.warp_tree_nodes));   // bytes from smem_raw
    const float* s_light = reinterpret_cast<const float*>(s_table + 28);

    // per-frame constants of ray generation, RaycastKernel :11-16
    // Per-frame constants of ray generation (RaycastKernel :11-16) arrive precomputed in the parameter block (wm1 = w-1,
    // hm1 = h-1, aspect = w/h: single IEEE operations, identical on the host; tan(fov/2) from the device, see csg_tan_kernel).
    // With supersampling (ss samples per axis) they describe the virtual (width*ss) x (height*ss) grid: sub-sample (sx,sy)
    // of pixel (x,y) is virtual pixel (x*ss+sx, y*ss+sy), SURVEY.md §8(d) row 5.
    const int ss = kSuper ? p.ss : 1;

    // ---- phase 1: macro tiles entirely outside the screen-space bound of the scene are Miss everywhere (:109): fill them
    // with the background, statically partitioned over the warps of this shard (no traversal, no tickets).
    {
        const int warps_per_cta = kThreads / 32;
        const int gw = blockIdx.x * warps_per_cta + (tid >> 5), GW = gridDim.x * warps_per_cta;
        const int total_macros = p.macro_x * p.macro_y;
        // fill_stride/fill_first: a single GPU or the root of a sharded frame fills every background tile itself (local
        // stores); the other shards fill none — only traced pixels cross NVLink
        for (int m = gw * p.fill_stride + p.fill_first; m < total_macros; m += GW * p.fill_stride) {
            const int my = p.div_magic ? (int)__umulhi((unsigned int)m, p.div_magic) : m / p.macro_x;
            const int mx = m - my * p.macro_x;
            if (my < p.band_m0 || my >= p.band_m1) continue;                                                   // another band of this frame
            if (mx >= p.rm_x0 && mx < p.rm_x0 + p.rm_w && my >= p.rm_y0 && my < p.rm_y0 + p.rm_h) continue;   // traced in phase 2
            const int x0 = mx * kMacroW, y0 = my * kMacroH;
            if (MODE == OUT_RGBA8 && (p.width & 3) == 0) {
                const uint32_t bg = 20u | (20u << 8) | (28u << 16) | 0xFF000000u;   // (0.08,0.08,0.11,1) quantised per Q12
                const int x = x0 + (lane & 15) * 4;
#pragma unroll 4
                for (int r = lane >> 4; r < kMacroH; r += 2) {
                    const int y = y0 + r;
                    if (x < p.width && y < p.height) reinterpret_cast<uint4*>(p.out)[((size_t)y * p.width + x) >> 2] = make_uint4(bg, bg, bg, bg);
                }
            } else {
                for (int r = 0; r < kMacroH; ++r) {
                    const int y = y0 + r;
                    if (y >= p.height) break;
                    for (int cx = lane; cx < kMacroW; cx += 32) {
                        const int x = x0 + cx;
                        if (x >= p.width) continue;
                        const size_t pix = (size_t)y * p.width + x;
                        if (MODE == OUT_RGBA8) reinterpret_cast<uint32_t*>(p.out)[pix] = 20u | (20u << 8) | (28u << 16) | 0xFF000000u;
                        else if (MODE == OUT_F32) reinterpret_cast<float4*>(p.out)[pix] = make_float4(0.08f, 0.08f, 0.11f, 1.0f);
                        else {
                            if (p.aov_hit) p.aov_hit[pix] = 0;
                            if (p.aov_prim) p.aov_prim[pix] = -1;
                            if (p.aov_t) p.aov_t[pix] = -1.0f;
                            if (p.aov_iters) p.aov_iters[pix] = 0;
                        }
                    }
                }
            }
        }
    }

    // everything below reads what csg_prune_kernel wrote (tile descriptors, pruned trees, hand-out order)
    cudaGridDependencySynchronize();

    // ---- phase 2: dynamic tile scheduling over the macro tiles that touch the bound: one ticket per warp tile; the next
    // ticket is requested before the current tile is rendered so the atomic's round trip hides behind the traversal.
    // Ticket t -> macro tile number (t >> 6) * shard_count + shard_rank of the rm_w x rm_h macro rectangle, warp tile t & 63.
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(p.tile_counter, 1u) - p.counter_base;
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    // Supersampling with 4 or 16 rays per pixel (sp = log2 of that): the samples of a pixel sit in neighbouring lanes instead
    // of being looped over by one lane: a warp pass covers 8 / 2 pixels of one row with one ray per lane, and a warp tile is
    // 4 / 16 such passes, handed out in tickets of 1 << gp passes each — so a heavy pixel does not serialise 16 traversals in
    // one warp, and the per-ticket set-up (descriptor, tree copy) is still shared by a few passes.
    const int sp = kSuper ? p.sp_shift : 0, gp = kSuper ? p.sp_group : 0;
    while (ticket < (unsigned int)p.n_local_warp_tiles) {
        const unsigned int cur = ticket >> (sp - gp);
        const int pass0 = (int)(ticket & ((1u << (sp - gp)) - 1u)) << gp;
        unsigned int next = 0;
        if (lane == 0) next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
        const int k = (int)(cur & 63u);
        // traced tiles are handed out heaviest first (order[] from csg_prune_kernel: tiles whose pruned tree is larger come
        // first, so that the expensive tiles are not the ones still running when the ticket counter runs dry)
        const int tile_no = p.order ? (int)__ldg(p.order + (cur >> 6)) : (int)(cur >> 6);
        const int j = tile_no * p.shard_count + p.shard_rank;
        // j / rm_w by multiply-high with a host-computed reciprocal (exact for the ranges the host enables it for)
        const int jy = p.rm_magic ? (int)__umulhi((unsigned int)j, p.rm_magic) : j / p.rm_w;
        const int mx = p.rm_x0 + (j - jy * p.rm_w), my = p.rm_y0 + jy;
        const int kx = (k & 1) | ((k >> 1) & 2) | ((k >> 2) & 4);        // Morton order inside the macro tile
        const int ky = ((k >> 1) & 1) | ((k >> 2) & 2) | ((k >> 3) & 4);
        const int tx0 = mx * kMacroW + kx * kWarpTileW, ty0 = my * kMacroH + ky * kWarpTileH;   // the warp tile's corner
        ticket = 0xffffffffu;   // placeholder; the real value is broadcast at the end of the iteration
        if (tx0 >= p.width || ty0 >= p.height) { ticket = __shfl_sync(0xffffffffu, next, 0); continue; }

        // this macro tile's pruned tree (csg_prune_kernel): only the primitives its rays can reach, operators whose other
        // operand cannot be reached collapsed away.  n_nodes == 0: every ray of the tile is a Miss.
        const int slot = p.shard_shift >= 0 ? (my * p.macro_x + mx) >> p.shard_shift : (my * p.macro_x + mx) / p.shard_count;
        const uint4 td = p.desc ? __ldg(reinterpret_cast<const uint4*>(p.desc) + slot) : make_uint4(0u, (uint32_t)p.n_nodes, p.full_flags, 0u);
        const unsigned char* tree = reinterpret_cast<const unsigned char*>(p.pool + 2 * (size_t)td.x);
        if (td.y != 0u && td.y <= (uint32_t)p.warp_tree_nodes) {
            uint4* my_tree = reinterpret_cast<uint4*>(smem_raw + my_tree_off);
            __syncwarp();   // everybody is done with the previous tile's copy
            for (uint32_t i = lane; i < 2u * td.y; i += 32u) my_tree[i] = __ldg(p.pool + 2 * (size_t)td.x + i);
            __syncwarp();
            tree = reinterpret_cast<const unsigned char*>(my_tree);
        }
        // whole warp tile outside the screen-space bound of the root box: every ray is a Miss (:109 background colour)
        const bool tile_empty = td.y == 0u || tx0 > p.rect_x1 || tx0 + (kWarpTileW - 1) < p.rect_x0 || ty0 > p.rect_y1 || ty0 + (kWarpTileH - 1) < p.rect_y0;

#pragma unroll 1
        for (int pass = pass0; pass < pass0 + (1 << gp); ++pass) {
            int x = tx0, y = ty0;   // this lane's pixel
            if (sp == 0) { x += lane & 7; y += lane >> 3; }
            else {   // 32 >> sp pixels of one row per pass, 1 << sp lanes per pixel
                const int ppw = 32 >> sp, lg = sp > 2 ? sp - 2 : 0;   // 8 / ppw = 1 << lg passes per row of the warp tile
                x += (pass & ((1 << lg) - 1)) * ppw + (lane >> sp);
                y += pass >> lg;
            }
            const bool active = x < p.width && y < p.height;
            const uint32_t pix = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;   // :33 (csg_upload keeps width*height below 2^31)
            if (__ballot_sync(0xffffffffu, active) == 0u) continue;

            Hit res = make_miss();
            int iters = 0;
            Ray r;
            r.ox = ox; r.oy = oy; r.oz = oz;
            float accx = 0.f, accy = 0.f, accz = 0.f;
            if (tile_empty) {
                const float w = (float)(ss * ss);
                accx = 0.08f * w; accy = 0.08f * w; accz = 0.11f * w;
            } else if (active) {
                const int n_samples = sp ? 1 : ss * ss;
#pragma unroll 1
                for (int s = 0, sx = sp ? ((lane & ((1 << sp) - 1)) & (ss - 1)) : 0, sy = sp ? ((lane & ((1 << sp) - 1)) >> (sp >> 1)) : 0; s < n_samples; ++s) {
                    const int vx = x * ss + sx, vy = y * ss + sy;
                    if (++sx == ss) { sx = 0; ++sy; }
                    // RaycastKernel :11-27 + Ray ctor (Ray.cuh:12-18)
                    const float u = __fdiv_rn(__fadd_rn((float)vx, 0.5f), p.wm1);
                    const float v = __fdiv_rn(__fadd_rn((float)vy, 0.5f), p.hm1);
                    const float nx = __fmul_rn(__fmul_rn(p.aspect, __fmaf_rn(u, 2.0f, -1.0f)), p.tan_half_fov);
                    const float ny = __fmul_rn(__fsub_rn(1.0f, __fadd_rn(v, v)), p.tan_half_fov);
                    float cx = __fadd_rn(p.forward[0], __fmaf_rn(p.right[0], nx, __fmul_rn(p.up[0], ny)));
                    float cy = __fadd_rn(p.forward[1], __fmaf_rn(p.right[1], nx, __fmul_rn(p.up[1], ny)));
                    float cz = __fadd_rn(p.forward[2], __fmaf_rn(p.right[2], nx, __fmul_rn(p.up[2], ny)));
#pragma unroll
                    for (int rep = 0; rep < 2; ++rep) {   // normalize() then the Ray ctor normalises again (Q3)
                        const float inv = __frcp_rn(__fsqrt_rn(dot_ref(cx, cy, cz, cx, cy, cz)));
                        cx = __fmul_rn(inv, cx); cy = __fmul_rn(inv, cy); cz = __fmul_rn(inv, cz);
                    }
                    r.dx = cx; r.dy = cy; r.dz = cz;
                    r.ix = rcp_approx(cx); r.iy = rcp_approx(cy); r.iz = rcp_approx(cz);   // culling boxes only: they carry 1e-5 of slack
                    res = traverse<MODE == OUT_AOV>(tree, p.prims, s_table, my_stack, (uint32_t)(kThreads * sizeof(uint4)), r, (td.z & kTileRootLeaf) != 0u,
                                                    (td.z & kTileRootPure) != 0u, p.root_is_leaf == 0, iters);
                    if (MODE != OUT_AOV) {
                        const float4 c = shade_pixel(res, r, p.prims, p, s_light);
                        accx += c.x; accy += c.y; accz += c.z;
                    }
                }
            }

            if (MODE == OUT_AOV) {
                if (active) {
                    const bool hit = !is_miss(res);
                    if (p.aov_hit) p.aov_hit[pix] = hit ? 1 : 0;
                    if (p.aov_prim) p.aov_prim[pix] = hit ? (int32_t)((res.m & H_META_MASK) >> H_ID_SHIFT) : -1;
                    if (p.aov_t) p.aov_t[pix] = hit ? res.t : -1.0f;
                    if (p.aov_iters) p.aov_iters[pix] = iters;
                }
            } else {
                if (kSuper && sp && !tile_empty) {
                    // sum the pixel's samples in sample order (the order of the one-lane loop), in every lane of the pixel
                    const int base = lane & ~((1 << sp) - 1);
                    float sxr = 0.f, syr = 0.f, szr = 0.f;
                    for (int s = 0; s < (1 << sp); ++s) {
                        sxr += __shfl_sync(0xffffffffu, accx, base + s);
                        syr += __shfl_sync(0xffffffffu, accy, base + s);
                        szr += __shfl_sync(0xffffffffu, accz, base + s);
                    }
                    accx = sxr; accy = syr; accz = szr;
                }
                float4 c = make_float4(accx, accy, accz, 1.0f);
                if (ss > 1) {   // box filter of the linear colour
                    const float w = __frcp_rn((float)(ss * ss));
                    c.x *= w; c.y *= w; c.z *= w;
                }
                if (kSuper && sp) {   // one lane per pixel stores
                    if (active && (lane & ((1 << sp) - 1)) == 0) {
                        if (MODE == OUT_F32) reinterpret_cast<float4*>(p.out)[pix] = c;
                        else reinterpret_cast<uint32_t*>(p.out)[pix] = to_u8(c.x) | (to_u8(c.y) << 8) | (to_u8(c.z) << 16) | 0xFF000000u;
                    }
                } else if (MODE == OUT_F32) {
                    if (active) reinterpret_cast<float4*>(p.out)[pix] = c;
                } else {
                    const uint32_t px8 = to_u8(c.x) | (to_u8(c.y) << 8) | (to_u8(c.z) << 16) | 0xFF000000u;
                    // four horizontally adjacent pixels -> one 16-byte store
                    const uint32_t p1 = __shfl_down_sync(0xffffffffu, px8, 1);
                    const uint32_t p2 = __shfl_down_sync(0xffffffffu, px8, 2);
                    const uint32_t p3 = __shfl_down_sync(0xffffffffu, px8, 3);
                    if ((p.width & 3) == 0) {
                        if (active && (lane & 3) == 0) reinterpret_cast<uint4*>(p.out)[pix >> 2] = make_uint4(px8, p1, p2, p3);
                    } else if (active) {
                        reinterpret_cast<uint32_t*>(p.out)[pix] = px8;
                    }
                }
            }
        }
        ticket = __shfl_sync(0xffffffffu, next, 0);
    }
}

// tan(cam.fov / 2.0f) of RaycastKernel :15-16, evaluated with the device tanf once per field of view (kept out of
// the frame kernel: tanf's large-argument path needs a local-memory scratch array).
__global__ void csg_tan_kernel(float fov, float* out) { *out = tanf(__fmul_rn(fov, 0.5f)); }

// ---- per-tile tree pruning ---------------------------------------------------------------------------------------------
// One CTA per traced macro tile builds the tile's own tree: a primitive whose culling box lies outside the tile's frustum
// (the 64x32 pixels plus a margin of one pixel) is a Miss for every ray of the tile, whatever tmin; an operator with such an
// operand behaves exactly like its other operand (Union: [x][M] -> RetL, [M][x] -> RetR; Difference: [x][M] -> RetL) or is a
// Miss itself (Difference without its left operand, Intersection without either: all M* cells, RaycastingKernels.cu:666-677),
// without ever looping — so dropping those primitives and collapsing those operators changes no result.  The surviving
// nodes are written in preorder, origin-relative (the subtractions of isBVHNodeHit :724-729, cubeHit :389-394 and
// sphereHit :139-143 are done here once per node and tile instead of once per ray: same single FADD, same bits), with
// operator boxes recomputed over what is left.
__device__ __forceinline__ void stage_record(const uint4 ua, const uint4 ub, float ox, float oy, float oz, uint4& oa, uint4& ob)
{
    float4 a = as_float4(ua), b = as_float4(ub);
    if ((ub.w & 7u) == 3u) {
        a.x = __fsub_rn(ox, a.x); a.y = __fsub_rn(oy, a.y); a.z = __fsub_rn(oz, a.z);
    } else {
        a.x = __fsub_rn(a.x, ox); a.y = __fsub_rn(a.y, oy); a.z = __fsub_rn(a.z, oz);
        a.w = __fsub_rn(a.w, ox); b.x = __fsub_rn(b.x, oy); b.y = __fsub_rn(b.y, oz);
    }
    oa = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
    ob = make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), ub.z, ub.w);
}

// culling box of a node relative to the origin: operators, cubes, cylinders carry it; spheres: centre +- r, padded like the host does
__device__ __forceinline__ void rel_cull_box(const uint4 ua, const uint4 ub, float ox, float oy, float oz, float lo[3], float hi[3])
{
    const float4 a = as_float4(ua), b = as_float4(ub);
    if ((ub.w & 7u) == 3u) {
        const float r = fabsf(a.w);
        const float c[3] = {a.x, a.y, a.z}, o[3] = {ox, oy, oz};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pad = r * 1e-4f + fabsf(c[k]) * 4e-7f + 1e-30f;
            lo[k] = (c[k] - r - pad) - o[k];
            hi[k] = (c[k] + r + pad) - o[k];
        }
    } else {
        lo[0] = a.x - ox; lo[1] = a.y - oy; lo[2] = a.z - oz;
        hi[0] = a.w - ox; hi[1] = b.x - oy; hi[2] = b.y - oz;
    }
}

constexpr int kPruneWarps = 4, kPruneThreads = kPruneWarps * 32;
constexpr int kListMax = 512;     // nodes one tile may look at (alive nodes + their tested children)
constexpr int kSlotMax = 256;     // records per tile slot
constexpr int kLevelMax = 64;
constexpr int kCostBuckets = 64;

struct PruneWarpSmem {            // working set of one warp = one tile
    int lnode[kListMax];          // node id, in breadth-first order of discovery
    short lchild[kListMax];       // operators: list position of the left child (the right child follows it)
    short lrep[kListMax];         // list position of the node standing for this subtree: itself, a descendant, or -1
    short lsize[kListMax];        // survivors in the subtree (valid where lrep[p] == p)
    short lidx[kListMax];         // preorder index among the survivors
    unsigned char lkind[kListMax];   // kind | alive << 3 | reachable << 4
    unsigned char flg[kListMax];  // bit0 pure, bit1 bounded
    float box[kListMax][6];       // culling box, origin-relative (operators: recomputed over what survives)
    short lvl[kLevelMax + 2];     // list position where each level starts
};

// One WARP per traced macro tile, top-down: only nodes whose parent is reachable from the tile are ever looked at, so the
// cost follows the size of the tile's own tree, not of the scene.  Three passes over the levels of the visited part:
// down (frustum tests), up (which operators survive, their boxes), down (preorder numbering and emission).
__global__ void __launch_bounds__(kPruneThreads) csg_prune_kernel(const __grid_constant__ PruneParams q)
{
    extern __shared__ __align__(16) unsigned char psm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float ox = q.cam_pos[0], oy = q.cam_pos[1], oz = q.cam_pos[2];
    const int N = q.n_nodes, S = q.slot_nodes;
    const int tile_ctas = (q.n_tiles + kPruneWarps - 1) / kPruneWarps;
    cudaTriggerProgrammaticLaunchCompletion();   // the frame kernel may be scheduled now; it waits for this grid before it reads our output

    if ((int)blockIdx.x >= tile_ctas) {
        // staging CTAs: origin-relative copy of the whole tree at the head of the pool, for tiles whose tree does not fit a slot
        const int nb = (int)gridDim.x - tile_ctas;
        for (int i = ((int)blockIdx.x - tile_ctas) * kPruneThreads + tid; i < N; i += nb * kPruneThreads) {
            uint4 oa, ob;
            stage_record(__ldg(&q.nodes[2 * i]), __ldg(&q.nodes[2 * i + 1]), ox, oy, oz, oa, ob);
            q.pool[2 * i] = oa;
            q.pool[2 * i + 1] = ob;
        }
        return;
    }
    const int tile = (int)blockIdx.x * kPruneWarps + warp;
    if (tile >= q.n_tiles) return;
    {   // pull the tree into L2 in one go (the walk below touches it level by level, one dependent miss at a time otherwise)
        const char* base = reinterpret_cast<const char*>(q.nodes);
        const size_t bytes = (size_t)N * 32;
        for (size_t off = ((size_t)(tile & 7) * 32 + lane) * 128; off < bytes && off < (size_t)(1 << 20); off += 8 * 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
    PruneWarpSmem& w = reinterpret_cast<PruneWarpSmem*>(psm)[warp];
    const unsigned int lt = (1u << lane) - 1u;

    // ---- the tile and its frustum
    const int j = tile * q.shard_count + q.shard_rank;
    const int jy = q.rm_magic ? (int)__umulhi((unsigned int)j, q.rm_magic) : j / q.rm_w;
    const int mx = q.rm_x0 + (j - jy * q.rm_w), my = q.rm_y0 + jy;
    const int slot = (my * q.macro_x + mx) / q.shard_count;
    float pn[5][3];   // inward plane normals: 4 sides through the origin + the camera plane
    {
        const float x0 = (float)(mx * kMacroW - 1) * q.ss, x1 = (float)(min(mx * kMacroW + kMacroW, q.width) + 1) * q.ss;
        const float y0 = (float)(my * kMacroH - 1) * q.ss, y1 = (float)(min(my * kMacroH + kMacroH, q.height) + 1) * q.ss;
        float d[4][3];
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // corner rays (RaycastKernel :11-25, un-normalised), around the tile: (x0,y0) (x1,y0) (x1,y1) (x0,y1)
            const float fx = (c == 1 || c == 2) ? x1 : x0, fy = (c >= 2) ? y1 : y0;
            const float u = fx / q.wm1, v = fy / q.hm1;
            const float nx = q.aspect * (2.0f * u - 1.0f) * q.tan_half_fov, ny = (1.0f - 2.0f * v) * q.tan_half_fov;
#pragma unroll
            for (int k = 0; k < 3; ++k) d[c][k] = q.forward[k] + q.right[k] * nx + q.up[k] * ny;
        }
        float dc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) dc[k] = d[0][k] + d[1][k] + d[2][k] + d[3][k];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float* a = d[c];
            const float* b = d[(c + 1) & 3];
            float n0 = a[1] * b[2] - a[2] * b[1], n1 = a[2] * b[0] - a[0] * b[2], n2 = a[0] * b[1] - a[1] * b[0];
            if (n0 * dc[0] + n1 * dc[1] + n2 * dc[2] < 0.0f) { n0 = -n0; n1 = -n1; n2 = -n2; }
            pn[c][0] = n0; pn[c][1] = n1; pn[c][2] = n2;
        }
        pn[4][0] = q.forward[0]; pn[4][1] = q.forward[1]; pn[4][2] = q.forward[2];
    }

    // ---- A. breadth-first from the root, queueing the operands of live operators.  Two ways to decide "live":
    //   pass 0 (frustum walk): test every visited node's box against the frustum.  Cost follows the number of boxes the
    //          frustum touches — small when the tree is spatially coherent.
    //   pass 1 (leaf marks; taken when pass 0 overflows its list, or first for small trees): test all primitives once, mark
    //          the way from every reachable primitive up to the root (2 bits per node: left / right operand has something
    //          below), then walk down along the marks only.  An operator with one marked side either stands for that side
    //          (Union; Difference when it is the left one) or is gone (Difference without its left operand, Intersection) and
    //          is skipped on the spot, so the list holds only operators with both sides marked, and primitives.
    uint32_t* mk = reinterpret_cast<uint32_t*>(psm + kPruneWarps * sizeof(PruneWarpSmem)) + (size_t)warp * q.mark_words;
    int total = 1, levels = 0;
    bool overflow = true;
    for (int pass = q.marks_first ? 1 : 0; pass < 2 && overflow; ++pass) {
        if (pass == 1) {
            if (q.mark_words == 0) break;           // tree too large for the marks
            for (int i = lane; i < q.mark_words; i += 32) mk[i] = 0u;
            __syncwarp();
            // all primitives once, coalesced and several loads in flight: culling box (world space) + node number, 32 B each
#pragma unroll 4
            for (int k = lane; k < q.n_leaves; k += 32) {
                const float4 la = __ldg(&q.leaf_boxes[2 * k]), lb4 = __ldg(&q.leaf_boxes[2 * k + 1]);
                const float lo[3] = {la.x - ox, la.y - oy, la.z - oz}, hi[3] = {lb4.x - ox, lb4.y - oy, lb4.z - oz};
                bool outside = false;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const float m = fmaxf(pn[c][0] * lo[0], pn[c][0] * hi[0]) + fmaxf(pn[c][1] * lo[1], pn[c][1] * hi[1]) +
                                    fmaxf(pn[c][2] * lo[2], pn[c][2] * hi[2]);
                    outside = outside || (m < 0.0f);
                }
                if (outside) continue;
                const int i = __float_as_int(la.w);
                atomicOr(&mk[i >> 4], 1u << ((i & 15) * 2));
                int c = i, par = __ldg(&q.parent[i]);
                while (par >= 0) {                   // up to the root, or to a node somebody else already marked
                    const int sh = (par & 15) * 2;
                    const uint32_t old = atomicOr(&mk[par >> 4], (c == par + 1 ? 1u : 2u) << sh);
                    if ((old >> sh) & 3u) break;
                    c = par; par = __ldg(&q.parent[par]);
                }
            }
            __syncwarp();
        }
        if (lane == 0) { w.lnode[0] = 0; w.lvl[0] = 0; }
        __syncwarp();
        int lb = 0, le = 1;
        total = 1; levels = 0; overflow = false;
        while (lb < le && !overflow) {
            if (lane == 0) w.lvl[levels + 1] = (short)le;
            int next_total = total;
            for (int base = lb; base < le; base += 32) {
                const int p = base + lane;
                const bool have = p < le;
                bool grow = false;
                uint32_t meta = 0u;
                int n = 0;
                if (have) {
                    n = w.lnode[p];
                    bool outside = false;
                    uint4 ua, ub;
                    if (pass == 1) {
                        for (;;) {
                            meta = __ldg(&q.nodes[2 * n + 1]).w;
                            const uint32_t kind = meta & 7u, m = (mk[n >> 4] >> ((n & 15) * 2)) & 3u;
                            if (kind >= 3u) { outside = !(m & 1u); break; }
                            if (m == 3u) break;
                            if (m == 1u && kind != 2u) { n = n + 1; continue; }              // stands for its left operand
                            if (m == 2u && kind == 0u) { n = (int)(meta >> 8); continue; }   // Union: stands for its right operand
                            outside = true;
                            break;
                        }
                        w.lnode[p] = n;
                    }
                    ua = __ldg(&q.nodes[2 * n]); ub = __ldg(&q.nodes[2 * n + 1]);
                    meta = ub.w;
                    float lo[3], hi[3];
                    rel_cull_box(ua, ub, ox, oy, oz, lo, hi);
                    if (pass == 0) {
#pragma unroll
                        for (int c = 0; c < 5; ++c) {
                            const float m = fmaxf(pn[c][0] * lo[0], pn[c][0] * hi[0]) + fmaxf(pn[c][1] * lo[1], pn[c][1] * hi[1]) +
                                            fmaxf(pn[c][2] * lo[2], pn[c][2] * hi[2]);
                            outside = outside || (m < 0.0f);
                        }
                    }
                    const uint32_t kind = meta & 7u;
                    w.lkind[p] = (unsigned char)(kind | (outside ? 0u : 8u));
                    w.lrep[p] = outside ? (short)-1 : (short)p;
                    w.lsize[p] = 1;
                    w.flg[p] = (unsigned char)(((kind == 3u || kind == 5u) ? 1u : 0u) | ((kind != 4u) ? 2u : 0u));
#pragma unroll
                    for (int c = 0; c < 3; ++c) { w.box[p][c] = lo[c]; w.box[p][3 + c] = hi[c]; }
                    grow = !outside && kind < 3u;
                }
                const unsigned int mask = __ballot_sync(0xffffffffu, grow);
                const int add = 2 * __popc(mask);
                if (next_total + add > kListMax) { overflow = true; break; }
                if (grow) {
                    const int c = next_total + 2 * __popc(mask & lt);
                    w.lnode[c] = n + 1;
                    w.lnode[c + 1] = (int)(meta >> 8);
                    w.lchild[p] = (short)c;
                }
                next_total += add;
            }
            __syncwarp();
            lb = le; le = next_total; total = next_total;
            if (++levels >= kLevelMax) overflow = true;
        }
    }
    uint32_t kept = 0u, flags = 0u;
    uint4* dst = q.pool + 2 * ((size_t)q.slots_off32 + (size_t)slot * S);
    int r0 = -1;
    if (!overflow) {
        // ---- B. deepest level first: an operator stays (both operands matter: box over what is left, flags), collapses to
        //         one operand, or goes
        for (int d = levels - 1; d >= 0; --d) {
            for (int p = w.lvl[d] + lane; p < w.lvl[d + 1]; p += 32) {
                const uint32_t k = w.lkind[p];
                if (!(k & 8u) || (k & 7u) >= 3u) continue;
                const int c = w.lchild[p];
                const int a = w.lrep[c], b = w.lrep[c + 1];
                const uint32_t kind = k & 7u;
                const int rp = kind == 0u ? (a < 0 ? b : (b < 0 ? a : p)) : kind == 1u ? (a < 0 ? -1 : (b < 0 ? a : p)) : ((a < 0 || b < 0) ? -1 : p);
                w.lrep[p] = (short)rp;
                if (rp != p) continue;
                w.lsize[p] = (short)(1 + w.lsize[a] + w.lsize[b]);
                const float* bl = w.box[a];
                const float* br = w.box[b];
                float* bo = w.box[p];
                if (kind == 0u) {                   // Union: both operands
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) { bo[c2] = fminf(bl[c2], br[c2]); bo[3 + c2] = fmaxf(bl[3 + c2], br[3 + c2]); }
                } else if (kind == 1u) {            // Difference: a subset of the left operand
#pragma unroll
                    for (int c2 = 0; c2 < 6; ++c2) bo[c2] = bl[c2];
                } else {                            // Intersection: a subset of both; the smaller box
                    float vl = 1.f, vr = 1.f;
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) { vl *= fmaxf(bl[3 + c2] - bl[c2], 0.f); vr *= fmaxf(br[3 + c2] - br[c2], 0.f); }
                    const float* bs = vl <= vr ? bl : br;
#pragma unroll
                    for (int c2 = 0; c2 < 6; ++c2) bo[c2] = bs[c2];
                }
                const uint32_t fl = w.flg[a], fr = w.flg[b];
                w.flg[p] = (unsigned char)(((kind == 0u) ? (fl & fr & 1u) : 0u) | (fl & fr & 2u));
            }
            __syncwarp();
        }
        r0 = w.lrep[0];
        if (r0 >= 0) {
            kept = (uint32_t)w.lsize[r0];
            if (kept > (uint32_t)S) overflow = true;
        }
    }
    if (!overflow && r0 >= 0) {
        // ---- C. root first: preorder index of every survivor (left operand right after its operator, right operand after the
        //         left subtree), and its record
        if (lane == 0) { w.lidx[r0] = 0; w.lkind[r0] |= 16u; }
        __syncwarp();
        for (int d = 0; d < levels; ++d) {
            for (int p = w.lvl[d] + lane; p < w.lvl[d + 1]; p += 32) {
                const uint32_t k = w.lkind[p];
                if (!(k & 16u)) continue;
                const int i = w.lidx[p];
                if ((k & 7u) >= 3u) {           // primitive: origin-relative record
                    const int n = w.lnode[p];
                    uint4 oa, ob;
                    stage_record(__ldg(&q.nodes[2 * n]), __ldg(&q.nodes[2 * n + 1]), ox, oy, oz, oa, ob);
                    dst[2 * i] = oa;
                    dst[2 * i + 1] = ob;
                    continue;
                }
                const int c = w.lchild[p];
                const int ra = w.lrep[c], rb = w.lrep[c + 1];
                const int r = i + 1 + w.lsize[ra];
                w.lidx[ra] = (short)(i + 1);
                w.lidx[rb] = (short)r;
                w.lkind[ra] |= 16u;
                w.lkind[rb] |= 16u;
                const uint32_t f = w.flg[p];
                const uint32_t meta = (k & 7u) | ((uint32_t)r << 8) | ((w.lkind[ra] & 7u) >= 3u ? kMetaLeftLeaf : 0u) |
                                      ((w.lkind[rb] & 7u) >= 3u ? kMetaRightLeaf : 0u) | ((f & 2u) ? kMetaBounded : 0u) | ((f & 1u) ? kMetaPure : 0u);
                const float* bo = w.box[p];
                dst[2 * i] = make_uint4(__float_as_uint(bo[0]), __float_as_uint(bo[1]), __float_as_uint(bo[2]), __float_as_uint(bo[3]));
                dst[2 * i + 1] = make_uint4(__float_as_uint(bo[4]), __float_as_uint(bo[5]), 0u, meta);
            }
            __syncwarp();
        }
        const uint32_t rk = w.lkind[r0] & 7u;
        flags = (rk >= 3u ? kTileRootLeaf : 0u) | ((rk < 3u && (w.flg[r0] & 1u)) ? kTileRootPure : 0u);
    }
    // ---- descriptor; heavy tiles (more nodes) are handed out first by the frame kernel: bucket lists, then one ordered list
    const uint32_t cost = overflow ? (uint32_t)min(N, 2 * kCostBuckets - 1) : kept;
    const int bucket = (int)min(cost / 2u, (uint32_t)(kCostBuckets - 1));
    if (lane == 0) {
        q.desc[slot] = overflow ? TileDesc{0u, (uint32_t)N, q.full_flags, 0u}
                                : TileDesc{kept ? q.slots_off32 + (uint32_t)slot * (uint32_t)S : 0u, kept, flags, 0u};
        if (q.order) {
            const unsigned int rank = atomicAdd(&q.hist[bucket], 1u);
            q.lists[(size_t)bucket * q.n_slots + rank] = (unsigned short)tile;
            __threadfence();
        }
    }
    if (!q.order) return;
    unsigned int done = 0;
    if (lane == 0) done = atomicAdd(q.done, 1u);
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done != (unsigned int)q.n_tiles - 1u) return;
    // last warp of the grid: concatenate the bucket lists, heaviest bucket first, and reset the counters for the next frame
    __threadfence();
    static_assert(kCostBuckets == 64, "two buckets per lane");
    unsigned int* start = reinterpret_cast<unsigned int*>(&w);   // the warp's own scratch is free now
    {
        // k = 63 - bucket: heaviest bucket first.  lane l owns k = l and k = 32 + l; start[k] = first position of bucket k in order[]
        const unsigned int c0 = __ldcg(&q.hist[63 - lane]), c1 = __ldcg(&q.hist[31 - lane]);
        unsigned int i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += t0; i1 += t1; }
        }
        const unsigned int first_half = __shfl_sync(0xffffffffu, i0, 31);
        start[lane] = i0 - c0;
        start[32 + lane] = first_half + i1 - c1;
    }
    __syncwarp();
    for (int b = lane; b < kCostBuckets; b += 32) q.hist[b] = 0u;
    if (lane == 0) *q.done = 0u;
    // every output position looks up its bucket (largest k with start[k] <= i): independent loads, several in flight per lane
#pragma unroll 4
    for (int i = lane; i < q.n_tiles; i += 32) {
        int k = 0;
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1)
            if (k + step < kCostBuckets && start[k + step] <= (unsigned int)i) k += step;
        q.order[i] = __ldcg(q.lists + (size_t)(63 - k) * q.n_slots + ((unsigned int)i - start[k]));
    }
}

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

struct Shard {  // one GPU's share of the frame
    int device = 0;
    int rank = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr;
    uint4* d_nodes = nullptr;
    uint4* d_pool = nullptr;     // [staged whole tree][one slot of slot_nodes records per macro tile of this shard]
    TileDesc* d_desc = nullptr;  // per macro tile of this shard
    int* d_parent = nullptr;
    float4* d_leaf_boxes = nullptr;      // per primitive: culling box (world space) + node number, 2 x float4
    unsigned int* d_hist = nullptr;      // kCostBuckets counters + 1 "done" counter
    unsigned short* d_lists = nullptr;   // kCostBuckets x n_slots
    unsigned short* d_order = nullptr;   // n_slots
    int n_slots = 0;
    float4* d_prims = nullptr;
    unsigned int* d_counter = nullptr;
    unsigned int counter_base = 0;
    float* d_tan = nullptr;
    int grid = 0;
    int n_local_warp_tiles = 0;
    uint8_t* target = nullptr;   // where this shard writes RGBA8 (root framebuffer, possibly a peer pointer)
    void* ipc_mapped = nullptr;
};

}  // namespace

struct csg_context {
    int width = 0, height = 0;
    int macro_x = 0, macro_y = 0;
    int shard_count = 1;
    bool multi_process = false;
    FlatTree tree;
    bool prune = true;           // per-tile pruned trees (csg_prune_kernel); off for trees too large for its shared memory
    int slot_nodes = 0;          // records per tile slot
    size_t prune_smem = 0;
    uint32_t full_flags = 0;
    int mark_words = 0, marks_first = 0;
    csg_scene scene_copy;        // for the second frame slot of csg_render_batch
    csg_context* twin = nullptr; // second frame slot (own stream, trees, framebuffer), created on first use
    cudaEvent_t ev_batch0 = nullptr, ev_batch1 = nullptr;
    int warp_tree_nodes = 0;     // per-warp shared-memory copy of the current tile's tree: capacity in records
    int band_m0 = 0, band_m1 = 0;   // macro-tile rows the next enqueue covers (0, 0 = the whole frame)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_band[8] = {};
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
    cudaError_t e = c->ss > 1 ? cudaLaunchKernelEx(&cfg, csg_frame_kernel<MODE, T, true>, fp)
                              : cudaLaunchKernelEx(&cfg, csg_frame_kernel<MODE, T, false>, fp);
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
    CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaFuncSetAttribute(csg_frame_kernel<MODE, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, csg_frame_kernel<MODE, T, true>, T, smem));
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

void fill_params(const csg_context* c, const Shard& s, const csg_camera* cam, const float light[3], FrameParams& fp)
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
    // background tiles: the shard that owns the framebuffer fills all of them; with an external target everybody fills its own
    const bool owns_fb = s.rank == 0 && !c->external_target;
    fp.fill_stride = owns_fb ? 1 : c->shard_count;
    fp.fill_first = owns_fb ? 0 : (c->external_target ? s.rank : (1 << 30));
    fp.shard_shift = -1;
    for (int b = 0; b < 16; ++b) if ((1 << b) == c->shard_count) fp.shard_shift = b;
    fp.counter_base = s.counter_base;
    fp.tile_counter = s.d_counter;
    fp.pool = s.d_pool;
    fp.desc = c->prune ? s.d_desc : nullptr;
    fp.order = c->prune ? s.d_order : nullptr;
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
    const long long traced = (long long)fp.rm_w * fp.rm_h;
    const long long mine = traced > s.rank ? (traced - s.rank + c->shard_count - 1) / c->shard_count : 0;
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
    fp.n_local_warp_tiles = (int)((mine * 64) << (fp.sp_shift - fp.sp_group));
    fp.rm_magic = (fp.rm_w > 1 && fp.rm_w < 4096 && traced < (1ll << 20)) ? (unsigned int)((1ull << 32) / (unsigned long long)fp.rm_w + 1ull) : 0u;
    if (fp.rm_w == 0) fp.rm_w = 1;   // never divide by zero; n_local_warp_tiles is 0 anyway
}

// Enqueue one frame on every shard.  mode: OUT_RGBA8 -> out = rgba8 target (NULL: each shard's own target),
// OUT_F32 / OUT_AOV only on single-shard contexts.
int enqueue_frame(csg_context* c, const csg_camera* cam, const float light[3], int mode, void* out)
{
    if (!c || !cam) return fail(CSG_ERR_ARG, "null argument");
    if (mode != OUT_RGBA8 && c->shards.size() != 1) return fail(CSG_ERR_ARG, "f32/aov output needs a single-shard context");
    if (mode == OUT_AOV && c->ss != 1) return fail(CSG_ERR_ARG, "AOV output is per primary ray: set supersampling to 1");
    Shard& root = c->shards[0];
    CU(cudaSetDevice(root.device));
    if (!(cam->fov == c->cached_fov)) {   // new field of view: one tiny launch + 4-byte readback, then cached
        csg_tan_kernel<<<1, 1, 0, root.stream>>>(cam->fov, root.d_tan);
        CU(cudaMemcpyAsync(&c->cached_tan, root.d_tan, sizeof(float), cudaMemcpyDeviceToHost, root.stream));
        CU(cudaStreamSynchronize(root.stream));
        c->cached_fov = cam->fov;
        c->launches++;
    }
    if (c->band_m0 == 0) CU(cudaEventRecord(root.ev_start, root.stream));   // later bands of a banded frame keep the first band's start mark
    for (size_t i = 1; i < c->shards.size(); ++i) {   // peers start after the root's start mark
        CU(cudaSetDevice(c->shards[i].device));
        CU(cudaStreamWaitEvent(c->shards[i].stream, root.ev_start, 0));
    }
    const int total_warps_per_cta = c->threads / 32;
    for (Shard& s : c->shards) {
        CU(cudaSetDevice(s.device));
        FrameParams fp;
        fill_params(c, s, cam, light, fp);
        c->last_rm[0] = fp.rm_x0; c->last_rm[1] = fp.rm_y0; c->last_rm[2] = fp.n_local_warp_tiles ? fp.rm_w : 0; c->last_rm[3] = fp.n_local_warp_tiles ? fp.rm_h : 0;
        {   // per-tile pruned trees + the staged copy of the whole tree for this camera
            PruneParams q;
            std::memset(&q, 0, sizeof q);
            for (int i = 0; i < 3; ++i) {
                q.cam_pos[i] = fp.cam_pos[i]; q.forward[i] = fp.forward[i]; q.right[i] = fp.right[i]; q.up[i] = fp.up[i];
            }
            q.tan_half_fov = fp.tan_half_fov; q.wm1 = fp.wm1; q.hm1 = fp.hm1; q.aspect = fp.aspect; q.ss = fp.ss;
            q.width = fp.width; q.height = fp.height; q.macro_x = fp.macro_x;
            q.rm_x0 = fp.rm_x0; q.rm_y0 = fp.rm_y0; q.rm_w = fp.rm_w; q.rm_magic = fp.rm_magic;
            q.shard_rank = fp.shard_rank; q.shard_count = fp.shard_count;
            q.n_tiles = c->prune ? (fp.n_local_warp_tiles >> (fp.sp_shift - fp.sp_group)) / 64 : 0;
            q.nodes = s.d_nodes; q.n_nodes = fp.n_nodes; q.parent = s.d_parent;
            q.leaf_boxes = s.d_leaf_boxes; q.n_leaves = (int)(c->tree.leaf_boxes.size() / 8);
            q.mark_words = c->mark_words; q.marks_first = c->marks_first;
            q.n_slots = s.n_slots; q.hist = s.d_hist; q.done = s.d_hist ? s.d_hist + kCostBuckets : nullptr;
            q.lists = s.d_lists; q.order = s.d_order;
            q.pool = s.d_pool; q.desc = s.d_desc; q.slot_nodes = c->slot_nodes;
            q.slots_off32 = (uint32_t)fp.n_nodes; q.full_flags = c->full_flags;
            const int stage_ctas = std::max(1, std::min(64, (fp.n_nodes + kPruneThreads - 1) / kPruneThreads));
            csg_prune_kernel<<<(q.n_tiles + kPruneWarps - 1) / kPruneWarps + stage_ctas, kPruneThreads, c->prune_smem, s.stream>>>(q);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return fail(CSG_ERR_CUDA, std::string("prune kernel launch: ") + cudaGetErrorString(e));
            c->launches++;
        }
        int rc = CSG_OK;
        if (mode == OUT_RGBA8) {
            fp.out = out ? out : (void*)s.target;
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
    if (scene->scene.nodes.empty()) return fail(CSG_ERR_ARG, "empty scene");
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
    c->stack_levels = std::max(1, c->tree.depth);
    c->root_box_valid = c->tree.root_box_valid;
    for (int i = 0; i < 6; ++i) c->root_box[i] = c->tree.root_box[i];

    auto cleanup_fail = [&](int code) { csg_free_context(c); return code; };

    // shared memory plan: [table 128 B][stack 16 B x (levels+2) x threads][tree 32 B/node]
    CU(cudaSetDevice(devices[0]));
    int max_optin = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, devices[0]));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devices[0]));
    const size_t tree_bytes = c->tree.nodes.size() * sizeof(NodeRec);
    const size_t table_bytes = 32 * sizeof(uint32_t);
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
        c->prune = !(off && off[0] == '1');
        c->prune_alloc = c->prune;
    }
    {
        // resident warps per SM for every (shape, tree placement); +1 KB per CTA is what the driver reserves
        const size_t sm_total = (size_t)max_optin + 1024;
        int best_warps = -1;
        const char* force_shape = std::getenv("CSG_B200_SHAPE");   // tuning aid: force a CTA shape
        for (int i = 0; i < kShapes; ++i) {
            const int T = kShapeThreads[i];
            if (force_shape && std::atoi(force_shape) != T) continue;
            const size_t base_need = (size_t)(c->stack_levels + 2) * T * sizeof(uint4) + table_bytes;   // +2: sentinel frame, search marker
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
        CU(cudaSetDevice(s.device));
        int bps = 0, rc;
        rc = configure_shape(c->threads, c->smem_bytes, &bps);
        if (rc) return cleanup_fail(rc);
        if (bps < 1) { g_err = "kernel does not fit on an SM"; return cleanup_fail(CSG_ERR_LIMIT); }
        int dev_sms = 0;
        CU(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, s.device));
        const int my_macros = (total_macros - s.rank + shard_count - 1) / shard_count;
        s.n_local_warp_tiles = my_macros * 64;
        const int want = (s.n_local_warp_tiles + (c->threads / 32) - 1) / (c->threads / 32);
        s.grid = std::max(1, std::min(dev_sms * bps, want));   // persistent CTAs: a multiple of the SM count
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&s.ev_start));
        CU(cudaEventCreate(&s.ev_done));
        CU(cudaMalloc(&s.d_nodes, std::max<size_t>(tree_bytes, 32)));
        CU(cudaMemcpy(s.d_nodes, c->tree.nodes.data(), tree_bytes, cudaMemcpyHostToDevice));
        {
            s.n_slots = my_macros;
            const size_t pool_records = c->tree.nodes.size() + (c->prune ? (size_t)s.n_slots * c->slot_nodes : 0);
            CU(cudaMalloc(&s.d_pool, std::max<size_t>(pool_records, 1) * sizeof(NodeRec)));
            CU(cudaMalloc(&s.d_desc, std::max<size_t>(s.n_slots, 1) * sizeof(TileDesc)));
            CU(cudaMemset(s.d_desc, 0, std::max<size_t>(s.n_slots, 1) * sizeof(TileDesc)));
            CU(cudaMalloc(&s.d_parent, std::max<size_t>(c->tree.parent.size(), 1) * sizeof(int)));
            CU(cudaMemcpy(s.d_parent, c->tree.parent.data(), c->tree.parent.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMalloc(&s.d_leaf_boxes, std::max<size_t>(c->tree.leaf_boxes.size(), 8) * sizeof(float)));
            CU(cudaMemcpy(s.d_leaf_boxes, c->tree.leaf_boxes.data(), c->tree.leaf_boxes.size() * sizeof(float), cudaMemcpyHostToDevice));
            if (c->prune && s.n_slots <= 65535) {   // tile numbers are stored as 16-bit
                CU(cudaMalloc(&s.d_hist, (kCostBuckets + 1) * sizeof(unsigned int)));
                CU(cudaMemset(s.d_hist, 0, (kCostBuckets + 1) * sizeof(unsigned int)));
                CU(cudaMalloc(&s.d_lists, (size_t)kCostBuckets * std::max(s.n_slots, 1) * sizeof(unsigned short)));
                CU(cudaMalloc(&s.d_order, (size_t)std::max(s.n_slots, 1) * sizeof(unsigned short)));
            }
            if (c->prune_smem > 48 * 1024)
                CU(cudaFuncSetAttribute(csg_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->prune_smem));
        }
        const size_t prim_bytes = c->tree.prims.size() * sizeof(PrimRec);
        CU(cudaMalloc(&s.d_prims, std::max<size_t>(prim_bytes, 80)));
        CU(cudaMemcpy(s.d_prims, c->tree.prims.data(), prim_bytes, cudaMemcpyHostToDevice));
        CU(cudaMalloc(&s.d_tan, sizeof(float)));
        CU(cudaMalloc(&s.d_counter, sizeof(unsigned int)));
        CU(cudaMemset(s.d_counter, 0, sizeof(unsigned int)));
        if (i == 0) {
            CU(cudaMalloc(&c->d_fb, (size_t)width * height * 4));
            CU(cudaMemset(c->d_fb, 0, (size_t)width * height * 4));
        } else {
            // NVLink peer stores into the root framebuffer
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, s.device, devices[0]));
            if (!can) { g_err = "device " + std::to_string(s.device) + " cannot access device " + std::to_string(devices[0]) + " (P2P)"; return cleanup_fail(CSG_ERR_CUDA); }
            cudaError_t pe = cudaDeviceEnablePeerAccess(devices[0], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { g_err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe); return cleanup_fail(CSG_ERR_CUDA); }
            cudaGetLastError();
        }
        s.target = c->d_fb;
    }
    char buf[512];
    std::snprintf(buf, sizeof buf,
                  "{\"threads_per_cta\": %d, \"ctas\": %d, \"sms\": %d, \"smem_bytes_per_cta\": %zu, \"tree_bytes\": %zu, "
                  "\"prune\": %s, \"slot_nodes\": %d, \"stack_levels\": %d, \"n_nodes\": %zu, \"n_prims\": %zu, \"shards\": %d, "
                  "\"macro_tiles\": %d, \"optimize\": %d}",
                  c->threads, c->shards[0].grid, sms, c->smem_bytes, tree_bytes, c->prune ? "true" : "false", c->slot_nodes,
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

size_t csg_scene_write(const csg_scene* scene, char* buf, size_t buflen)
{
    if (!scene) return 0;
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
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
        for (cudaEvent_t e : c->ev_band) if (e) cudaEventDestroy(e);
    }
    if (c->ev_batch0) cudaEventDestroy(c->ev_batch0);
    if (c->ev_batch1) cudaEventDestroy(c->ev_batch1);
    for (Shard& s : c->shards) {
        cudaSetDevice(s.device);
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.ipc_mapped) cudaIpcCloseMemHandle(s.ipc_mapped);
        cudaFree(s.d_nodes);
        cudaFree(s.d_pool);
        cudaFree(s.d_desc);
        cudaFree(s.d_parent);
        cudaFree(s.d_leaf_boxes);
        cudaFree(s.d_hist);
        cudaFree(s.d_lists);
        cudaFree(s.d_order);
        cudaFree(s.d_prims);
        cudaFree(s.d_counter);
        cudaFree(s.d_tan);
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
    return CSG_OK;
}

int csg_set_gather_target(csg_context* ctx, uint8_t* rgba8_dev)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    for (Shard& s : ctx->shards) s.target = rgba8_dev ? rgba8_dev : ctx->d_fb;
    ctx->external_target = rgba8_dev != nullptr;
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

int csg_render(csg_context* ctx, const csg_camera* cam, const csg_light* light, uint8_t* rgba8_out)
{
    if (!ctx || !cam || !light || !rgba8_out) return fail(CSG_ERR_ARG, "null argument");
    const bool dev = is_device_pointer(rgba8_out);
    float ld[3];
    csg_light_direction(light, ld);
    // multi-shard contexts always gather into the root framebuffer first
    const bool direct = dev && ctx->shards.size() == 1;
    Shard& root = ctx->shards[0];
    const size_t bytes = (size_t)ctx->width * ctx->height * 4;
    int n_bands = (!dev && ctx->shards.size() == 1 && ctx->shard_count == 1 && !ctx->external_target && ctx->macro_y >= 16) ? 4 : 1;
    if (const char* nb = std::getenv("CSG_B200_BANDS")) n_bands = n_bands > 1 ? std::min(std::max(std::atoi(nb), 1), 8) : 1;   // tuning aid
    if (n_bands > 1) {
        // Host output on one GPU: the frame is rendered in 4 bands of macro-tile rows; band k travels over PCIe (the 33 MB copy
        // is 3x the render time at 4K) while band k+1 renders.
        CU(cudaSetDevice(root.device));
        if (!ctx->copy_stream) {
            CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (cudaEvent_t& e : ctx->ev_band) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        int rc = CSG_OK;
        int rm[4] = {0, 0, 0, 0};   // union of the bands' traced rectangles, for csg_prune_stats
        for (int b = 0; b < n_bands && !rc; ++b) {
            ctx->band_m0 = ctx->macro_y * b / n_bands;
            ctx->band_m1 = ctx->macro_y * (b + 1) / n_bands;
            rc = enqueue_frame(ctx, cam, ld, OUT_RGBA8, nullptr);
            if (rc) break;
            if (ctx->last_rm[2] > 0 && ctx->last_rm[3] > 0) {
                if (rm[3] == 0) { rm[0] = ctx->last_rm[0]; rm[1] = ctx->last_rm[1]; rm[2] = ctx->last_rm[2]; }
                rm[3] = ctx->last_rm[1] + ctx->last_rm[3] - rm[1];
            }
            const size_t r0 = (size_t)ctx->band_m0 * kMacroH, r1 = std::min<size_t>((size_t)ctx->band_m1 * kMacroH, (size_t)ctx->height);
            const size_t off = r0 * ctx->width * 4, len = (r1 - r0) * ctx->width * 4;
            CU(cudaEventRecord(ctx->ev_band[b], root.stream));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_band[b], 0));
            CU(cudaMemcpyAsync(rgba8_out + off, ctx->d_fb + off, len, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        ctx->band_m0 = ctx->band_m1 = 0;
        for (int i = 0; i < 4; ++i) ctx->last_rm[i] = rm[i];
        if (rc) return rc;
        CU(cudaStreamSynchronize(ctx->copy_stream));
        return sync_frame(ctx);
    }
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

int csg_set_pruning(csg_context* ctx, int enabled)
{
    if (!ctx) return fail(CSG_ERR_ARG, "null context");
    ctx->prune = enabled != 0 && ctx->prune_alloc;
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
            if (j % ctx->shard_count != s.rank) continue;
            const int m = (ctx->last_rm[1] + jy) * ctx->macro_x + ctx->last_rm[0] + jx;
            const TileDesc& t = d[(size_t)(m / ctx->shard_count)];
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
    Shard& root = ctx->shards[0];
    CU(cudaSetDevice(root.device));
    csg_tan_kernel<<<1, 1, 0, root.stream>>>(fov, root.d_tan);
    CU(cudaMemcpyAsync(out, root.d_tan, sizeof(float), cudaMemcpyDeviceToHost, root.stream));
    CU(cudaStreamSynchronize(root.stream));
    ctx->launches++;
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

const char* csg_context_info(csg_context* ctx) { return ctx ? ctx->info.c_str() : ""; }

}  // extern "C"
