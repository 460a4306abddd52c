// csg_frame.cuh — csg_frame_kernel: ray generation, traversal of the tile's tree, hit details, Phong, framebuffer write.
// Device side of Raycaster::Raycast (RayCasting/Raycaster.cu:23-34): RaycastKernel + LightningKernel
// (RayCasting/Kernels/RaycastingKernels.cu:3-111) fused into one persistent kernel.  Included by csg_render.cu only.
#pragma once
#include "csg_kernel.cuh"

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
constexpr int kWarpTreeMax = 64;                // records of the largest per-warp shared-memory tree copy
constexpr int kSmemHead = 384;                  // bytes before the stack: outcome table + light (128 B), first positions of the 64 cost buckets (256 B); csg_render.cu sizes the launch with it

// Hit details + Phong (sphere/cylinder/cubeHitDetails :183-200/:338-372/:436-457 and LightningKernel :49-111).
// (Inlined: every kernel has one call site, and out of line the call cost ~20 instructions of argument moves per shaded warp tile:
// frame kernel alone 0.1155 -> 0.1136 ms, configs[4] 7.19 -> 7.11 ms.)
template <bool kCyl>
__device__ __forceinline__ float4 shade_pixel(const Hit res, const Ray r, const float4* __restrict__ prims, const FrameParams& p, const float* __restrict__ s_light)
{
    if (is_miss(res)) return make_float4(0.08f, 0.08f, 0.11f, 1.0f);   // :109
    const uint32_t id = (res.m & H_META_MASK) >> H_ID_SHIFT;
    const uint32_t kind = (res.m >> H_KIND_SHIFT) & 7u;
    const float4 col = __ldg(&prims[id * 5 + 0]);
    const float4 pc = __ldg(&prims[id * 5 + 1]);
    const float t = res.t;
    const float px = __fmaf_rn(t, r.dx, r.ox), py = __fmaf_rn(t, r.dy, r.oy), pz = __fmaf_rn(t, r.dz, r.oz);
    float nx = 0.0f, ny = 0.0f, nz = 0.0f;
    if (kind == 3u) {                                   // sphereHitDetails :189-195
        nx = px - pc.x; ny = py - pc.y; nz = pz - pc.z;
    } else if (kind == 5u) {                            // cubeHitDetails :446-451
        const float4 th = __ldg(&prims[id * 5 + 2]);
        nx = cube_normal_component(__fsub_rn(px, pc.x), pc.w, th.x, th.y);
        ny = cube_normal_component(__fsub_rn(py, pc.y), pc.w, th.x, th.y);
        nz = cube_normal_component(__fsub_rn(pz, pc.z), pc.w, th.x, th.y);
    } else if (kCyl) {                                  // cylinderHitDetails :345-365
        const float4 pb = __ldg(&prims[id * 5 + 2]);
        const float4 pv = __ldg(&prims[id * 5 + 3]);
        if (res.m & H_FLAG1) { nx = -pv.x; ny = -pv.y; nz = -pv.z; }
        else if (res.m & H_FLAG2) { nx = pv.x; ny = pv.y; nz = pv.z; }
        else {
            // Here the reference's kernel forms C = centre - (h/2)V with one rounding (FFMA -V, h/2, centre in its SASS), unlike
            // cylinderHit, whose C is the two-rounding base point of the primitive record
            const float hh = __fmul_rn(pb.w, 0.5f);
            const float cx = __fmaf_rn(-pv.x, hh, pc.x), cy = __fmaf_rn(-pv.y, hh, pc.y), cz = __fmaf_rn(-pv.z, hh, pc.z);
            const float ocx = r.ox - cx, ocy = r.oy - cy, ocz = r.oz - cz;
            const float dV = dot_ref(pv.x, pv.y, pv.z, r.dx, r.dy, r.dz);
            const float ocv = dot_ref(pv.x, pv.y, pv.z, ocx, ocy, ocz);
            const float m = __fmaf_rn(t, dV, ocv);
            nx = __fmaf_rn(-pv.x, m, px - cx);
            ny = __fmaf_rn(-pv.y, m, py - cy);
            nz = __fmaf_rn(-pv.z, m, pz - cz);
        }
    }
    // the normal (caps carry the unit axis as is; everything else is normalised) and the view vector of LightningKernel :78-103
    // are normalised side by side: one range check for both squared lengths, then two independent chains (inv_len_fast)
    const bool unit = (res.m & (H_FLAG1 | H_FLAG2)) && kind == 4u;
    float vx = p.cam_pos[0] - px, vy = p.cam_pos[1] - py, vz = p.cam_pos[2] - pz;
    const float nn = unit ? 1.0f : dot_ref(nx, ny, nz, nx, ny, nz), vv = dot_ref(vx, vy, vz, vx, vy, vz);
    float inv, iv;
    if (len2_safe(nn) && len2_safe(vv)) { inv = inv_len_fast(nn); iv = inv_len_fast(vv); }
    else { inv = inv_len_exact_path(nn); iv = inv_len_exact_path(vv); }
    if (!unit) { nx *= inv; ny *= inv; nz *= inv; }
    if (res.m & H_FLIP) { nx = -nx; ny = -ny; nz = -nz; }
    if ((res.m & H_CLS) == H_EXIT) { nx = -nx; ny = -ny; nz = -nz; }

    const float Lx = s_light[0], Ly = s_light[1], Lz = s_light[2];   // normalize(lightDir), once per CTA
    vx *= iv; vy *= iv; vz *= iv;
    const float in = inv_len(dot_ref(nx, ny, nz, nx, ny, nz));       // reflect() re-normalises n
    const float ux = nx * in, uy = ny * in, uz = nz * in;
    // dot(-L, n) of reflect(): the reference's SASS rounds the x product first here (FMUL Lx*nx; FFMA -Ly,ny,-that; FFMA -Lz,nz,.),
    // not the y product as in every other dot of the path
    const float dn = __fmaf_rn(-Lz, uz, __fmaf_rn(-Ly, uy, -__fmul_rn(Lx, ux)));
    const float two = dn + dn;
    const float rx = __fmaf_rn(-ux, two, -Lx), ry = __fmaf_rn(-uy, two, -Ly), rz = __fmaf_rn(-uz, two, -Lz);
    const float diff = fmaxf(dot_ref(nx, ny, nz, Lx, Ly, Lz), 0.0f);
    const float sb = fmaxf(dot_ref(vx, vy, vz, rx, ry, rz), 0.0f);
    const float spec = sb > 0.0f ? powf(sb, 30.0f) : 0.0f;   // powf(+0, 30) is exactly +0: skip the call for pixels facing away from the highlight
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
//   * flat operands (ST_FLAT): a Union over a few spheres is evaluated from the spheres' roots (eval_flat_union, csg_kernel.cuh).
// (Rounds 1-2 also had a nearest-Enter search for pure union subtrees — a closest-hit descent with a shrinking limit that fell back
// to the frame machine on an Exit or a tie.  With per-tile trees, sibling pruning and flat operands it had become a loss on every
// BASELINE config — Cheese512 0.1499 -> 0.1436 ms, configs[4] 8.91 -> 8.17 ms without it — and was removed; the pure flag of the
// tree records remains.)

template <bool COUNT, bool kCyl>
__device__ __forceinline__ Hit traverse(const unsigned char* __restrict__ tree, const float4* __restrict__ prims,
                                        const uint32_t* __restrict__ table, const uint32_t stack,
                                        const uint32_t stack_stride, const int& stack_levels, const Ray& r, const bool root_is_leaf,
                                        const bool root_gated, int& iters)
{
    enum { ST_ENTER = 0, ST_LOOPL = 1, ST_LOOPR = 2, ST_COMPUTE = 3, ST_RETURN = 4, ST_DONE = 5, ST_FLAT = 6 };
    Hit L = make_miss(), R = make_miss();      // ST_FLAT: R.t = the limit beyond which an Enter need not be found
    float tmin = 0.0f;                        // :466
    if (root_is_leaf) {
        // The scene's root is a primitive: GoTo's leaf branch on the virtual root, no box test (:582-594, Q7).
        // A pruned tile tree that collapsed to one primitive (root_gated): the primitive is still reached through its
        // operators in the reference, so a cylinder keeps its gating box (Q6).
        bool go;
        float tn;
        uint32_t cm;
        eval_child<kCyl>(tree, prims, 0u, r, tmin, root_gated, L, go, tn, cm);
        return L;
    }
    uint32_t n = 0u;                           // byte offset of the current operator's record
    uint32_t sp = stack;                       // next free frame (shared-memory address; stack_stride = bytes between levels)
    sts128(sp, make_uint4(0u, 0u, 0u, 0xffffffffu)); // sentinel frame: popping it ends the traversal (no base pointer to keep)
    sp += stack_stride;
    int st = ST_ENTER;
    if (*reinterpret_cast<const uint32_t*>(tree + 28) & kMetaFlat) { R.t = INFINITY; st = ST_FLAT; }   // the whole (tile) tree is one flat Union of spheres
    while (st != ST_DONE) {
        if (COUNT) iters += (st == ST_ENTER) ? (1 << 10) : 1;   // packed: (bits 20+: unused since the search is gone) | operator visits | other iterations
        if (st <= ST_LOOPR) {
            const uint32_t meta = *reinterpret_cast<const uint32_t*>(tree + n + 28);
            const uint32_t op = meta & 7u;
            const uint32_t cl = n + 32u, cr = (meta >> 8) << 5;
            Hit a = make_miss(), b = make_miss();
            bool goA = false, goB = false;
            float tnA = -INFINITY, tnB = -INFINITY;
            uint32_t mA = 0u, mB = 0u;
            if (st != ST_LOOPR) eval_child<kCyl>(tree, prims, cl, r, tmin, st == ST_ENTER, a, goA, tnA, mA);
            if (st == ST_ENTER && op != 0u && !goA && is_miss(a)) {
                // left operand of a Difference/Intersection already missed: the node's result is Miss whatever the right
                // operand does (all M* cells of both tables, :670-677) — skip the right subtree (Q8)
                L = a; R = a;
                st = ST_RETURN;
            } else {
                if (st != ST_LOOPL) eval_child<kCyl>(tree, prims, cr, r, tmin, st == ST_ENTER, b, goB, tnB, mB);
                if (st == ST_LOOPL) { L = a; st = ST_COMPUTE; }
                else if (st == ST_LOOPR) { R = b; st = ST_COMPUTE; }
                else {
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
                        float lim = INFINITY;
                        if (!goA) {                                                        // :556-561
                            sts128(sp, make_uint4(__float_as_uint(L.t), L.m | F_LOAD_LFT, 0u, n));
                            first = cr; fm = mB;
                            if (op != 2u && !is_miss(L)) lim = L.t;
                        } else if (!goB) {                                                 // :562-567
                            sts128(sp, make_uint4(__float_as_uint(R.t), R.m | F_LOAD_RGH, 0u, n));
                            first = cl; fm = mA;
                            if (op == 0u && !is_miss(R)) lim = R.t;
                        } else {                                                           // :568-574
                            const bool right_first = (op == 0u) && (tnB < tnA);
                            const uint32_t pend_flat = ((right_first ? mA : mB) >> 6) & 2u;   // bit 1: the pending operand is flat
                            sts128(sp, make_uint4(__float_as_uint(tmin), (right_first ? F_FIRST_RGH : F_FIRST_LFT) | pend_flat,
                                             __float_as_uint(right_first ? tnA : tnB), n));
                            first = right_first ? cr : cl; fm = right_first ? mB : mA;
                        }
                        sp += stack_stride; n = first;
                        if (fm & kMetaFlat) { R.t = lim; st = ST_FLAT; }   // R is free here: saved in the frame, or a Miss
                    }
                }
            }
        }
        if (st == ST_FLAT) {
            // (Behind the operator visit, so that a descent decided there is evaluated in the same round.)
            // n is a flat operand (a Union over a few spheres) the machine was about to descend into; the frame that takes its
            // result is pushed.  Its result follows from the spheres' roots (eval_flat_union; the frames above sp are free and hold its
            // list): return it as if the descent had happened — or, when eval_flat_union gives up, descend after all.
            const uint2 fe = eval_flat_union((uint32_t)__cvta_generic_to_shared(tree), n, r, tmin, R.t, sp, stack_stride, stack + (uint32_t)(stack_levels + 2) * stack_stride);
            if (fe.y != kFlatGaveUp) { L.t = __uint_as_float(fe.x); L.m = fe.y; R = L; st = ST_RETURN; }
            else st = ST_ENTER;
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
                        const float lim = (pop != 2u && !miss) ? L.t : INFINITY;
                        if (f.y & 2u) { R.t = lim; st = ST_FLAT; }
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
                else {   // into an operator: a flat one (word 6 of this record says so) is evaluated from its spheres' roots
                    sts128(sp, make_uint4(__float_as_uint(R.t), R.m | F_LOAD_RGH, 0u, n)); sp += stack_stride;
                    st = (*reinterpret_cast<const uint32_t*>(tree + n + 24) & kW6LeftFlat) ? ST_FLAT : ST_ENTER; n = n + 32u;
                    R.t = INFINITY;   // eval_flat_union's limit: none (R is saved in the frame)
                }
            } else if (o == O_LOOPR) {                                                 // :647-653
                tmin = R.t;
                if (meta & kMetaRightLeaf) st = ST_LOOPR;
                else {
                    sts128(sp, make_uint4(__float_as_uint(L.t), L.m | F_LOAD_LFT, 0u, n)); sp += stack_stride;
                    st = (*reinterpret_cast<const uint32_t*>(tree + n + 24) & kW6RightFlat) ? ST_FLAT : ST_ENTER; n = (meta >> 8) << 5;
                    R.t = INFINITY;   // eval_flat_union's limit: none (R is what gets re-evaluated)
                }
            } else { L = R = make_miss(); st = ST_RETURN; }                            // :654-660
            if (st == ST_RETURN && sp == stack + stack_stride) st = ST_DONE;           // only the sentinel is left: this is the root's result (L == R)
        }
    }
    return L;                                   // :511
}

#ifdef CSG_FRAME_PROBE   // one-off instrumented build (tools/gpu_frame_probe.py): per-warp timeline of the frame kernel
__device__ unsigned long long g_frame_probe[8192][8];
__device__ __forceinline__ unsigned long long probe_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define FPROBE(k, v) do { if (lane == 0 && pw < 8192) g_frame_probe[pw][k] = (v); } while (0)
#else
#define FPROBE(k, v) do { } while (0)
#endif

template <int MODE, int kThreads, bool kSuper, bool kCyl, bool kPair>   // kSuper: more than one ray per pixel (keeps the sample loop and its accumulators out of the common case); kCyl: the scene has cylinders; kPair: tickets of two warp tiles (one ray per pixel only)
__global__ void __launch_bounds__(kThreads, min_blocks_for(kThreads)) csg_frame_kernel(const __grid_constant__ FrameParams p)
{
    // shared memory: [outcome table 128 B][stack: (levels+2) x kThreads x 16 B][scratch frame: kThreads x 16 B][per-warp tree copy: warp_tree_nodes x 32 B].
    // Every 64x32-pixel macro tile has its own pruned, origin-relative tree (csg_prune_kernel), a few hundred bytes to a few
    // KB; a warp copies the tree of its current tile into shared memory when it fits, and reads it through L1 otherwise.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* s_table = reinterpret_cast<uint32_t*>(smem_raw);
    uint4* s_stack = reinterpret_cast<uint4*>(smem_raw + kSmemHead);

    const int tid = threadIdx.x, lane = tid & 31;
    const float ox = p.cam_pos[0], oy = p.cam_pos[1], oz = p.cam_pos[2];
#ifdef CSG_FRAME_PROBE
    const int pw = blockIdx.x * (kThreads / 32) + (tid >> 5);
    unsigned long long pr_longest = 0, pr_tiles = 0, pr_t0 = 0, pr_longest_ticket = 0;
    FPROBE(0, probe_now());
#endif

    if (tid < 27) s_table[tid] = kOutcomeTable[tid];
    if (kSuper && tid == 28) {
        // Background of a supersampled pixel none of whose samples hits: the miss colour (:109) summed ss*ss times in the order of
        // the sample loop (s_table[27] = sum of 0.08f, s_table[31] = sum of 0.11f) — in float that is not ss*ss times the colour
        // (sixteen 0.11f add up to 1.7600001), and the box filter of the reference's samples is what the frame must equal.
        float b8 = 0.0f, b11 = 0.0f;
        for (int i = 0; i < p.ss * p.ss; ++i) { b8 += 0.08f; b11 += 0.11f; }
        reinterpret_cast<float*>(s_table)[27] = b8;
        reinterpret_cast<float*>(s_table)[31] = b11;
    }
    if (tid == 27) {   // L = normalize(lightDir), LightningKernel :78: the same for every pixel of the frame
        const float il = __frcp_rn(__fsqrt_rn(dot_ref(p.light[0], p.light[1], p.light[2], p.light[0], p.light[1], p.light[2])));
        float* s_light = reinterpret_cast<float*>(s_table + 28);
        s_light[0] = p.light[0] * il; s_light[1] = p.light[1] * il; s_light[2] = p.light[2] * il;
    }
    __syncthreads();
    gate_enter(p.gate, 2);   // sharded frames: nothing of this frame happens before the root GPU has started it
    const uint32_t my_stack = (uint32_t)__cvta_generic_to_shared(s_stack + tid);   // frames are addressed in the shared window: 32-bit
    // per-warp copy of the current tile's tree (when it fits): traversal then reads shared memory instead of L1/L2
    // bytes from smem_raw to this warp's tree copy.  Worked out where it is used, once per tile, from a thread id the compiler
    // cannot hoist (volatile): kept across the traversal it is the one value too many for 80 registers (it was spilled)
    auto my_tree_off = [&p]() {
        uint32_t t;
        asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
        return (uint32_t)kSmemHead + 16u * (uint32_t)((p.stack_levels + 3) * kThreads + (int)(t >> 5) * (2 * p.warp_tree_nodes));
    };
    const float* s_light = reinterpret_cast<const float*>(s_table + 28);
    // supersampling: one more 16-byte frame per thread behind the traversal stack (colour accumulators)
    const uint32_t my_scratch = my_stack + (uint32_t)((p.stack_levels + 2) * kThreads) * 16u;

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
            const int my = p.div_magic ? (int)__umulhi((unsigned int)m, p.div_magic) : m;   // div_magic == 0: one macro tile per row
            const int mx = m - my * p.macro_x;
            if (my < p.band_m0 || my >= p.band_m1) continue;                                                   // another band of this frame
            if (p.shard_mode && my % p.shard_count != p.shard_rank) continue;                                  // another shard's row
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
                        else if (MODE == OUT_F32) {
                            if (kSuper) {   // the box filter of ss*ss miss samples (s_table[27] / [31])
                                const float bgw = __frcp_rn((float)(ss * ss));
                                const float bg8 = reinterpret_cast<const float*>(s_table)[27] * bgw, bg11 = reinterpret_cast<const float*>(s_table)[31] * bgw;
                                reinterpret_cast<float4*>(p.out)[pix] = make_float4(bg8, bg8, bg11, 1.0f);
                            } else reinterpret_cast<float4*>(p.out)[pix] = make_float4(0.08f, 0.08f, 0.11f, 1.0f);
                        }
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
    FPROBE(1, probe_now());
    cudaGridDependencySynchronize();
    FPROBE(2, probe_now());
    // Hand-out order of the traced tiles: heaviest first (tiles whose pruned tree is larger come first, so that the expensive tiles
    // are not the ones still running when the ticket counter runs dry).  The pruning kernel leaves one list of tile descriptors per
    // cost bucket and the bucket sizes; position m of the order is entry m - start[k] of bucket 63 - k, k the largest with
    // start[k] <= m, where start[] = exclusive prefix sums of the sizes, heaviest bucket first — worked out here by the first warp
    // of every CTA (one load, five shuffles) instead of an ordering pass by the last CTA of the pruning kernel, which the whole
    // frame used to wait 2-3 us for.
    unsigned int* s_start = reinterpret_cast<unsigned int*>(smem_raw + 128);
    // (the first ticket is asked for before the bucket sizes are read: the two round trips overlap)
    unsigned int ticket = 0;
    if (lane == 0) ticket = atomicAdd(p.tile_counter, 1u) - p.counter_base;
    if (p.lists) {
        // every warp works the prefix sums out (no divergent shuffles), the first one stores them
        const unsigned int c0 = __ldcg(&p.hist[63 - lane]), c1 = __ldcg(&p.hist[31 - lane]);   // lane l owns k = l and k = 32 + l
        unsigned int i0 = c0, i1 = c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
            if (lane >= o) { i0 += t0; i1 += t1; }
        }
        const unsigned int first_half = __shfl_sync(0xffffffffu, i0, 31);
        if (tid < 32) {
            s_start[lane] = i0 - c0;
            s_start[32 + lane] = first_half + i1 - c1;
        }
        __syncthreads();
    }
    // position m of the hand-out order -> the tile's descriptor (offset32, n_nodes, flags, tile number); called by whole warps
    auto ordered_tile = [&p, s_start, lane](unsigned int m) {
        const unsigned int k = (unsigned int)(__popc(__ballot_sync(0xffffffffu, s_start[lane] <= m)) + __popc(__ballot_sync(0xffffffffu, s_start[32 + lane] <= m))) - 1u;
        return __ldg(p.lists + (size_t)(63u - k) * (unsigned int)p.list_stride + (m - s_start[k]));
    };

    // ---- phase 2: dynamic tile scheduling over the macro tiles that touch the bound: one ticket per warp tile.  With one
    // ray per pixel the next ticket is requested when the traversal of the current tile is over, so that the atomic's round
    // trip hides behind the shading — not earlier: a ticket taken at the start of a heavy tile would sit with this warp while
    // others run dry (with few tiles per warp, e.g. a frame sharded over 8 GPUs, that decided the length of the frame).
    // With supersampling a ticket is a few passes of a warp tile and there are plenty: it is requested up front.
    // Ticket t -> macro tile number (t >> 6) * shard_count + shard_rank of the rm_w x rm_h macro rectangle, warp tile t & 63.
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    // Supersampling with 4 or 16 rays per pixel (sp = log2 of that): the samples of a pixel sit in neighbouring lanes instead
    // of being looped over by one lane: a warp pass covers 8 / 2 pixels of one row with one ray per lane, and a warp tile is
    // 4 / 16 such passes, handed out in tickets of 1 << gp passes each — so a heavy pixel does not serialise 16 traversals in
    // one warp, and the per-ticket set-up (descriptor, tree copy) is still shared by a few passes.
    const int sp = kSuper ? p.sp_shift : 0, gp = kSuper ? p.sp_group : 0;
    // RaycastKernel :11-27 + Ray ctor (Ray.cuh:12-18): direction of the ray through virtual pixel (vx, vy)
    // (x + 0.5) / (w - 1), (y + 0.5) / (h - 1): numerators in [0.5, 2^24]; with divisors in [1, 2^24] (uv_fast, uniform) the IEEE
    // quotients come from refined reciprocals of the two frame constants (div_shared, csg_kernel.cuh) — same bits as __fdiv_rn
    const bool uv_fast = p.wm1 >= 1.0f && p.wm1 <= 16777216.0f && p.hm1 >= 1.0f && p.hm1 <= 16777216.0f;
    auto make_ray = [&p, uv_fast](int vx, int vy, Ray& ray) {
        const float un = __fadd_rn((float)vx, 0.5f), vn = __fadd_rn((float)vy, 0.5f);
        float u, v;
        if (uv_fast) { u = div_shared(un, p.wm1, rcp_refined(p.wm1)); v = div_shared(vn, p.hm1, rcp_refined(p.hm1)); }
        else { u = __fdiv_rn(un, p.wm1); v = __fdiv_rn(vn, p.hm1); }
        const float nx = __fmul_rn(__fmul_rn(p.aspect, __fmaf_rn(u, 2.0f, -1.0f)), p.tan_half_fov);
        const float ny = __fmul_rn(__fsub_rn(1.0f, __fadd_rn(v, v)), p.tan_half_fov);
        float cx = __fadd_rn(p.forward[0], __fmaf_rn(p.right[0], nx, __fmul_rn(p.up[0], ny)));
        float cy = __fadd_rn(p.forward[1], __fmaf_rn(p.right[1], nx, __fmul_rn(p.up[1], ny)));
        float cz = __fadd_rn(p.forward[2], __fmaf_rn(p.right[2], nx, __fmul_rn(p.up[2], ny)));
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {   // normalize() then the Ray ctor normalises again (Q3)
            const float inv = inv_len(dot_ref(cx, cy, cz, cx, cy, cz));
            cx = __fmul_rn(inv, cx); cy = __fmul_rn(inv, cy); cz = __fmul_rn(inv, cz);
        }
        ray.dx = cx; ray.dy = cy; ray.dz = cz;
        ray.ix = rcp_approx(cx); ray.iy = rcp_approx(cy); ray.iz = rcp_approx(cz);   // culling boxes only: they carry 1e-5 of slack
    };
    while (ticket < (unsigned int)p.n_local_warp_tiles) {
#ifdef CSG_FRAME_PROBE
        if (pr_tiles == 0) FPROBE(3, probe_now());
        pr_t0 = probe_now();
        const unsigned int pr_ticket = ticket;
#endif
        if constexpr (!kSuper && kPair) {
            // ---- one ray per pixel, plenty of tickets (one GPU at 4K: csg_render.cu picks the kPair kernels when there are 12 warp
            // tiles per warp of the grid or more).  A ticket is TWO warp tiles, horizontal neighbours inside one macro tile (Morton
            // order: k and k + 1), so that the ticket's set-up — the atomic's and the ordered list's round trips, tile and tree
            // look-up, the tree copy — is paid once per 64 pixels.  When tickets are scarce (a frame sharded over several GPUs) the
            // fine hand-out decides the length of the frame: those launches take the one-tile path below.
            constexpr int ps = 1;
            const unsigned int macro = ticket >> (6 - ps);
            const int k = (int)((ticket << ps) & 63u);
            uint4 td = make_uint4(0u, (uint32_t)p.n_nodes, p.full_flags, 0u);
            int tile_no = (int)macro;
            if (p.lists) {   // heaviest first; an entry carries the tile's descriptor
                td = ordered_tile(macro);
                tile_no = (int)td.w;
            }
            int mx, my;
            shard_tile_coords(tile_no, p.shard_mode, p.shard_rank, p.shard_count, p.rm_x0, p.rm_y0, p.rm_w, p.rm_magic, p.row_first, mx, my);
            const int kx = (k & 1) | ((k >> 1) & 2) | ((k >> 2) & 4);        // Morton order inside the macro tile
            const int ky = ((k >> 1) & 1) | ((k >> 2) & 2) | ((k >> 3) & 4);
            const int tx0 = mx * kMacroW + kx * kWarpTileW, ty0 = my * kMacroH + ky * kWarpTileH;   // corner of the first warp tile
            ticket = 0xffffffffu;   // placeholder; the real value is broadcast at the end of the iteration
            if (!p.lists && p.desc) {   // natural order: the descriptor is looked up by position
                const int slot = slot_of_macro(p.shard_mode, mx, my, p.macro_x, p.shard_count);
                td = __ldg(reinterpret_cast<const uint4*>(p.desc) + slot);
            }
            // nothing to trace anywhere in the ticket: no primitive reachable from the macro tile, or all of it outside the frame /
            // the screen-space bound of the root box (every ray is a Miss, :109 background colour)
            const bool span_empty = td.y == 0u || tx0 >= p.width || ty0 >= p.height || tx0 > p.rect_x1 || tx0 + (kWarpTileW << ps) - 1 < p.rect_x0 ||
                                    ty0 > p.rect_y1 || ty0 + (kWarpTileH - 1) < p.rect_y0;
            const unsigned char* tree = reinterpret_cast<const unsigned char*>(p.pool + 2 * (size_t)td.x);
            const bool tree_fits = !span_empty && td.y <= (uint32_t)p.warp_tree_nodes;
            if (tree_fits) {
                // the tile's tree goes to this warp's shared-memory slice with asynchronous copies (no registers in between): up to
                // four 16-byte pieces per lane, in flight while the first ray is generated
                static_assert(kWarpTreeMax == 64, "four copies per lane cover 64 records");
                const uint32_t slice = (uint32_t)__cvta_generic_to_shared(smem_raw + my_tree_off());
                __syncwarp();   // everybody is done with the previous tile's copy
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint32_t i = (uint32_t)lane + 32u * q4;
                    if (i < 2u * td.y) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(slice + 16u * i), "l"(p.pool + 2 * (size_t)td.x + i) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                tree = smem_raw + my_tree_off();
            }
            unsigned int next = 0;
            int req_lane = -1;   // lane that has asked for the next ticket (-1: nobody yet)
#pragma unroll 1
            for (int half = 0; half < (1 << ps); ++half) {
                const int tx = tx0 + half * kWarpTileW;
                const int x = tx + (lane & 7), y = ty0 + (lane >> 3);   // this lane's pixel
                const bool active = x < p.width && y < p.height;
                const uint32_t pix = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;   // :33 (csg_upload keeps width*height below 2^31)
                const unsigned int amask = __ballot_sync(0xffffffffu, active);
                if (amask == 0u) continue;
                const bool tile_empty = span_empty || tx > p.rect_x1 || tx + (kWarpTileW - 1) < p.rect_x0;
                Hit res = make_miss();
                int iters = 0;
                float accx = 0.08f, accy = 0.08f, accz = 0.11f;
                if (!tile_empty) {
                    Ray r;
                    r.ox = ox; r.oy = oy; r.oz = oz;
                    make_ray(x, y, r);
                    if (tree_fits) {   // (nothing pending in the second half: the wait falls through)
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        __syncwarp();
                    }
                    if (active) {
                        res = traverse<MODE == OUT_AOV, kCyl>(tree, p.prims, s_table, my_stack, (uint32_t)(kThreads * sizeof(uint4)), p.stack_levels, r, (td.z & kTileRootLeaf) != 0u,
                                                              p.root_is_leaf == 0, iters);
                        // the last tile of the ticket is traced: ask for the next ticket now, so that the atomic's round trip hides behind
                        // the shading — not earlier: a ticket held during a heavy tile would sit with this warp while others run dry
                        if (half == (1 << ps) - 1 && lane == __ffs(amask) - 1) next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
                        if (MODE != OUT_AOV) {
                            const float4 c = shade_pixel<kCyl>(res, r, p.prims, p, s_light);
                            accx = c.x; accy = c.y; accz = c.z;
                        }
                    }
                    if (half == (1 << ps) - 1) req_lane = __ffs(amask) - 1;
                }
                if (MODE == OUT_AOV) {
                    if (active) {
                        const bool hit = !is_miss(res);
                        if (p.aov_hit) p.aov_hit[pix] = hit ? 1 : 0;
                        if (p.aov_prim) p.aov_prim[pix] = hit ? (int32_t)((res.m & H_META_MASK) >> H_ID_SHIFT) : -1;
                        if (p.aov_t) p.aov_t[pix] = hit ? res.t : -1.0f;
                        if (p.aov_iters) p.aov_iters[pix] = iters;
                    }
                } else if (MODE == OUT_F32) {
                    if (active) reinterpret_cast<float4*>(p.out)[pix] = make_float4(accx, accy, accz, 1.0f);
                } else {
                    const uint32_t px8 = to_u8(accx) | (to_u8(accy) << 8) | (to_u8(accz) << 16) | 0xFF000000u;
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
#ifdef CSG_FRAME_PROBE
            { const unsigned long long d = probe_now() - pr_t0; ++pr_tiles; if (d > pr_longest) { pr_longest = d; pr_longest_ticket = pr_ticket; } }
#endif
            if (req_lane < 0) {   // the last tile of the ticket was not traced (empty, or outside the frame)
                if (lane == 0) next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
                req_lane = 0;
            }
            ticket = __shfl_sync(0xffffffffu, next, req_lane);
            continue;
        }
        // ---- one warp tile per ticket, and supersampling (kSuper)
        const int ts = kSuper ? p.sp_tshift : 0;   // = sp - gp, precomputed: a plain parameter load is rematerialised, a difference kept alive across the tile spilled
        const unsigned int cur = ticket >> ts;
        const int pass0 = (int)(ticket & ((1u << ts) - 1u)) << gp;
        unsigned int next = 0;
        int req_lane = kSuper ? 0 : -1;   // lane that has asked for the next ticket (-1: nobody yet)
        if (kSuper) {   // requested up front; parked in the scratch frame (w) while the tile is traced, not in a register
            if (lane == 0) next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
            sts128(my_scratch, make_uint4(0u, 0u, 0u, next));
        }
        const int k = (int)(cur & 63u);
        // traced tiles are handed out heaviest first (order[] from the pruning kernel: tiles whose pruned tree is larger come
        // first, so that the expensive tiles are not the ones still running when the ticket counter runs dry); an entry carries
        // the tile's descriptor, so the tile and its tree are known after one load
        uint4 td = make_uint4(0u, (uint32_t)p.n_nodes, p.full_flags, 0u);
        int tile_no = (int)(cur >> 6);
        if (p.lists) {
            td = ordered_tile(cur >> 6);
            tile_no = (int)td.w;
        }
        int mx, my;
        shard_tile_coords(tile_no, p.shard_mode, p.shard_rank, p.shard_count, p.rm_x0, p.rm_y0, p.rm_w, p.rm_magic, p.row_first, mx, my);
        const int kx = (k & 1) | ((k >> 1) & 2) | ((k >> 2) & 4);        // Morton order inside the macro tile
        const int ky = ((k >> 1) & 1) | ((k >> 2) & 2) | ((k >> 3) & 4);
        const int tx0 = mx * kMacroW + kx * kWarpTileW, ty0 = my * kMacroH + ky * kWarpTileH;   // the warp tile's corner
        ticket = 0xffffffffu;   // placeholder; the real value is broadcast at the end of the iteration
        if (tx0 >= p.width || ty0 >= p.height) {
            if (!kSuper && lane == 0) next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
            ticket = __shfl_sync(0xffffffffu, next, 0);
            continue;
        }

        // this macro tile's pruned tree (csg_prune_kernel): only the primitives its rays can reach, operators whose other
        // operand cannot be reached collapsed away.  n_nodes == 0: every ray of the tile is a Miss.
        if (!p.lists && p.desc) {   // natural order: the descriptor is looked up by position
            const int slot = slot_of_macro(p.shard_mode, mx, my, p.macro_x, p.shard_count);
            td = __ldg(reinterpret_cast<const uint4*>(p.desc) + slot);
        }
        // whole warp tile outside the screen-space bound of the root box: every ray is a Miss (:109 background colour)
        const bool tile_empty = td.y == 0u || tx0 > p.rect_x1 || tx0 + (kWarpTileW - 1) < p.rect_x0 || ty0 > p.rect_y1 || ty0 + (kWarpTileH - 1) < p.rect_y0;
        const unsigned char* tree = reinterpret_cast<const unsigned char*>(p.pool + 2 * (size_t)td.x);
        const bool tree_fits = !tile_empty && td.y <= (uint32_t)p.warp_tree_nodes;
        Ray r0;   // one ray per pixel: this lane's ray, generated while the tree is on its way
        r0.ox = ox; r0.oy = oy; r0.oz = oz;
        r0.dx = r0.dy = r0.dz = r0.ix = r0.iy = r0.iz = 0.0f;
        if (!kSuper && !tile_empty) {
            // the records are requested first (up to four 16-byte loads per lane), the ray is generated while they are in flight
            // (~100 instructions, no memory), then they go to shared memory
            static_assert(kWarpTreeMax == 64, "four loads per lane cover 64 records");
            uint4 tr[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const uint32_t i = (uint32_t)lane + 32u * q4;
                tr[q4] = (tree_fits && i < 2u * td.y) ? __ldg(p.pool + 2 * (size_t)td.x + i) : make_uint4(0u, 0u, 0u, 0u);
            }
            make_ray(tx0 + (lane & 7), ty0 + (lane >> 3), r0);
            if (tree_fits) {
                uint4* my_tree = reinterpret_cast<uint4*>(smem_raw + my_tree_off());
                __syncwarp();   // everybody is done with the previous tile's copy
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint32_t i = (uint32_t)lane + 32u * q4;
                    if (i < 2u * td.y) my_tree[i] = tr[q4];
                }
                __syncwarp();
                tree = reinterpret_cast<const unsigned char*>(my_tree);
            }
        } else if (tree_fits) {
            uint4* my_tree = reinterpret_cast<uint4*>(smem_raw + my_tree_off());
            __syncwarp();   // everybody is done with the previous tile's copy
            for (uint32_t i = lane; i < 2u * td.y; i += 32u) my_tree[i] = __ldg(p.pool + 2 * (size_t)td.x + i);
            __syncwarp();
            tree = reinterpret_cast<const unsigned char*>(my_tree);
        }

        int pass = pass0;   // pass0 is a multiple of 1 << gp: the ticket's passes end where the low gp bits wrap
#pragma unroll 1
        do {
            int x = tx0, y = ty0;   // this lane's pixel
            if (sp == 0) { x += lane & 7; y += lane >> 3; }
            else {   // 32 >> sp pixels of one row per pass, 1 << sp lanes per pixel
                const int ppw = 32 >> sp, lg = sp > 2 ? sp - 2 : 0;   // 8 / ppw = 1 << lg passes per row of the warp tile
                x += (pass & ((1 << lg) - 1)) * ppw + (lane >> sp);
                y += pass >> lg;
            }
            const bool active = x < p.width && y < p.height;
            const uint32_t pix = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;   // :33 (csg_upload keeps width*height below 2^31)
            const unsigned int amask = __ballot_sync(0xffffffffu, active);
            if (amask == 0u) continue;

            Hit res = make_miss();
            int iters = 0;
            Ray r;
            r.ox = ox; r.oy = oy; r.oz = oz;
            float accx = 0.f, accy = 0.f, accz = 0.f;
            if (!kSuper && !tile_empty) req_lane = __ffs(amask) - 1;
            if (tile_empty) {
                accx = accy = kSuper ? reinterpret_cast<const float*>(s_table)[27] : 0.08f;
                accz = kSuper ? reinterpret_cast<const float*>(s_table)[31] : 0.11f;
            } else if (active) {
                if (kSuper) {
                    // More than one ray per pixel: the running colour sum lives in this thread's scratch frame in shared memory,
                    // not in registers, while a sample is traced (the traversal needs every register of the 80 there are; three
                    // accumulators and the sample counters across it were what spilled to local memory).
                    sts128(my_scratch, make_uint4(0u, 0u, 0u, lds128(my_scratch).w));
                    const int n_samples = sp ? 1 : ss * ss;
                    const int sl = lane & ((1 << sp) - 1);   // sample-parallel: this lane's sample of the pixel
#pragma unroll 1
                    for (int s = 0; s < n_samples; ++s) {
                        const int sy = sp ? (sl >> (sp >> 1)) : s / ss;
                        const int sx = sp ? (sl & (ss - 1)) : s - sy * ss;
                        make_ray(x * ss + sx, y * ss + sy, r);
                        res = traverse<MODE == OUT_AOV, kCyl>(tree, p.prims, s_table, my_stack, (uint32_t)(kThreads * sizeof(uint4)), p.stack_levels, r, (td.z & kTileRootLeaf) != 0u,
                                                        p.root_is_leaf == 0, iters);
                        if (MODE != OUT_AOV) {
                            const float4 c = shade_pixel<kCyl>(res, r, p.prims, p, s_light);
                            const uint4 a = lds128(my_scratch);
                            sts128(my_scratch, make_uint4(__float_as_uint(__uint_as_float(a.x) + c.x), __float_as_uint(__uint_as_float(a.y) + c.y),
                                                          __float_as_uint(__uint_as_float(a.z) + c.z), a.w));
                        }
                    }
                    const uint4 a = lds128(my_scratch);
                    accx = __uint_as_float(a.x); accy = __uint_as_float(a.y); accz = __uint_as_float(a.z);
                } else {
                    r.dx = r0.dx; r.dy = r0.dy; r.dz = r0.dz; r.ix = r0.ix; r.iy = r0.iy; r.iz = r0.iz;
                    res = traverse<MODE == OUT_AOV, kCyl>(tree, p.prims, s_table, my_stack, (uint32_t)(kThreads * sizeof(uint4)), p.stack_levels, r, (td.z & kTileRootLeaf) != 0u,
                                                    p.root_is_leaf == 0, iters);
                    if (lane == __ffs(amask) - 1)   // the tile is traced: ask for the next ticket now
                        next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
                    if (MODE != OUT_AOV) {
                        const float4 c = shade_pixel<kCyl>(res, r, p.prims, p, s_light);
                        accx = c.x; accy = c.y; accz = c.z;
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
#pragma unroll 1
                    for (int s = 0; s < (1 << sp); s += 4) {   // 4 or 16 samples: whole groups of four, no remainder loop
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            sxr += __shfl_sync(0xffffffffu, accx, base + s + u);
                            syr += __shfl_sync(0xffffffffu, accy, base + s + u);
                            szr += __shfl_sync(0xffffffffu, accz, base + s + u);
                        }
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
        } while ((++pass & ((1 << gp) - 1)) != 0);
#ifdef CSG_FRAME_PROBE
        { const unsigned long long d = probe_now() - pr_t0; ++pr_tiles; if (d > pr_longest) { pr_longest = d; pr_longest_ticket = pr_ticket; } }
#endif
        if (!kSuper && req_lane < 0) {   // nothing was traced (empty tile)
            if (lane == 0) next = atomicAdd(p.tile_counter, 1u) - p.counter_base;
            req_lane = 0;
        }
        if (kSuper) next = lds128(my_scratch).w;
        ticket = __shfl_sync(0xffffffffu, next, req_lane);
    }
#ifdef CSG_FRAME_PROBE
    FPROBE(4, probe_now()); FPROBE(5, pr_longest); FPROBE(6, pr_tiles); FPROBE(7, pr_longest_ticket);
#endif
    // ---- join of a sharded frame (GateParams): a peer's last CTA tells the root that all of this shard's pixels have landed in
    // the root's framebuffer; the root's last CTA waits for every peer before the kernel (and with it the frame) ends
    if (p.gate.role != GATE_NONE) {
        __syncthreads();
        if (tid == 0) {
            // behind the barrier one fence covers the pixel stores of the whole CTA (over NVLink on a peer): they are performed,
            // system-wide, before this CTA is counted
            if (p.gate.role == GATE_PEER) __threadfence_system();
            const unsigned int before = atomicAdd(p.gate.exit_counter, 1u);
            if (before == gridDim.x - 1u) {
                *p.gate.exit_counter = 0u;   // ready for the next frame
                SPROBE(4);
                // (release: ordered behind every CTA's count this thread has just observed, and with them behind their fenced stores)
                if (p.gate.role == GATE_PEER) st_release_sys(&p.gate.words->done[p.gate.rank], p.gate.seq);
                else
                    for (int r = 1; r < p.gate.n_shards; ++r) wait_seq(&p.gate.words->done[r], p.gate.seq, p.gate.err);
                SPROBE(5);
            }
        }
    }
}

// tan(cam.fov / 2.0f) of RaycastKernel :15-16, evaluated with the device tanf once per field of view (kept out of
// the frame kernel: tanf's large-argument path needs a local-memory scratch array).
__global__ void csg_tan_kernel(float fov, float* out) { *out = tanf(__fmul_rn(fov, 0.5f)); }

}  // namespace csgb
