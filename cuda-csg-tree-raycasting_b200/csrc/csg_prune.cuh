// csg_prune.cuh — csg_prune_kernel: one pruned, origin-relative tree per 64x32-pixel macro tile, rebuilt every frame.
// No counterpart in the reference (its per-frame work starts at RaycastKernel); see DESIGN.md 4.1 for why this cannot
// change a result.  Included by csg_render.cu only.
#pragma once
#include "csg_kernel.cuh"
#include "csg_frame.cuh"   // tile geometry constants

namespace csgb {

// ---- per-tile tree pruning ---------------------------------------------------------------------------------------------
// One warp per traced macro tile builds the tile's own tree: a primitive whose culling box lies outside the tile's frustum
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
        a.x = __fsub_rn(ox, a.x); a.y = __fsub_rn(oy, a.y); a.z = __fsub_rn(oz, a.z);   // sphereHit :139-143
        a.w = __fmaf_rn(a.w, a.w, -dot_ref(a.x, a.y, a.z, a.x, a.y, a.z));              // :146, r*r - dot(oc,oc): FFMA r,r,-dot in the reference's SASS
    } else {
        a.x = __fsub_rn(a.x, ox); a.y = __fsub_rn(a.y, oy); a.z = __fsub_rn(a.z, oz);
        a.w = __fsub_rn(a.w, ox); b.x = __fsub_rn(b.x, oy); b.y = __fsub_rn(b.y, oz);
    }
    uint32_t meta = ub.w;
    if ((meta & 7u) >= 4u && div_safe(a.x) && div_safe(a.y) && div_safe(a.z) && div_safe(a.w) && div_safe(b.x) && div_safe(b.y))
        meta |= kMetaBoxSafe;   // cube_isect / gate_box_exact may divide through shared reciprocals
    oa = make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w));
    ob = make_uint4(__float_as_uint(b.x), __float_as_uint(b.y), (ub.w & 7u) < 3u ? 0u : ub.z, meta);   // operators: no flat-evaluation flags (word 6, csg_scene.h)
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

// Traced macro tile number `tile` of this shard -> macro tile coordinates.
__device__ __forceinline__ void tile_of(const PruneParams& q, int tile, int& mx, int& my)
{
    shard_tile_coords(tile, q.shard_mode, q.shard_rank, q.shard_count, q.rm_x0, q.rm_y0, q.rm_w, q.rm_magic, q.row_first, mx, my);
}

// Frustum of a macro tile: its pixels plus a margin of one pixel.  Inward plane normals: 4 sides through the camera position
// (cross products of the un-normalised corner rays of RaycastKernel :11-25) + the camera plane.
__device__ __forceinline__ void tile_frustum(const PruneParams& q, int mx, int my, float pn[5][3])
{
    const float x0 = (float)(mx * kMacroW - 1) * q.ss, x1 = (float)(min(mx * kMacroW + kMacroW, q.width) + 1) * q.ss;
    const float y0 = (float)(my * kMacroH - 1) * q.ss, y1 = (float)(min(my * kMacroH + kMacroH, q.height) + 1) * q.ss;
    float d[4][3];
#pragma unroll
    for (int c = 0; c < 4; ++c) {   // corner rays around the tile: (x0,y0) (x1,y0) (x1,y1) (x0,y1)
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

// box [lo, hi] (relative to the camera position) entirely behind one of the five planes?
__device__ __forceinline__ bool box_outside(const float pn[5][3], const float lo[3], const float hi[3])
{
    bool outside = false;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const float m = fmaxf(pn[c][0] * lo[0], pn[c][0] * hi[0]) + fmaxf(pn[c][1] * lo[1], pn[c][1] * hi[1]) +
                        fmaxf(pn[c][2] * lo[2], pn[c][2] * hi[2]);
        outside = outside || (m < 0.0f);
    }
    return outside;
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
    gate_enter(q.gate);
    if (blockIdx.x == 0 && tid < kCostBuckets && q.hist_next) q.hist_next[tid] = 0u;   // the bucket counters of the NEXT pruned frame (this frame's: q.hist)

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
    int mx, my;
    tile_of(q, tile, mx, my);
    const int slot = slot_of_macro(q.shard_mode, mx, my, q.macro_x, q.shard_count);
    float pn[5][3];   // inward plane normals: 4 sides through the origin + the camera plane
    tile_frustum(q, mx, my, pn);

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
                if (box_outside(pn, lo, hi)) continue;
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
                    if (pass == 0) outside = box_outside(pn, lo, hi);
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
    // ---- descriptor; heavy tiles (more nodes) are handed out first by the frame kernel: one list per cost bucket, which the frame
    // kernel reads through the prefix sums of the bucket sizes (csg_frame.cuh) — no ordering pass behind the last tile
    const uint32_t cost = overflow ? (uint32_t)min(N, 2 * kCostBuckets - 1) : kept;
    const int bucket = (int)min(cost / 2u, (uint32_t)(kCostBuckets - 1));
    if (lane == 0) {
        const TileDesc td = overflow ? TileDesc{0u, (uint32_t)N, q.full_flags, 0u}
                                     : TileDesc{kept ? q.slots_off32 + (uint32_t)slot * (uint32_t)S : 0u, kept, flags, 0u};
        q.desc[slot] = td;
        if (q.lists) {   // the list entry carries the descriptor: the frame kernel gets the tile and its tree in one load
            const unsigned int rank = atomicAdd(&q.hist[bucket], 1u);
            q.lists[(size_t)bucket * q.n_slots + rank] = make_uint4(td.offset32, td.n_nodes, td.flags, (uint32_t)tile);
        }
    }
}

// ---- the same pruned trees, without walking the tree ----------------------------------------------------------------------
// csg_prune_kernel follows the tree level by level: ~12 dependent levels x 3 passes per tile, one warp per tile, every level a
// round trip to L2 — 14 us per tile however little work there is, which is the fixed part of a frame once it is spread over
// several GPUs.  csg_prune_flat_kernel (the default for trees of up to kFlatMaxNodes nodes) gets the same tree out of the
// PREORDER layout with prefix sums, one CTA per tile and no pointer chasing:
//   1. every primitive's culling box against the tile's frustum, in parallel (coalesced 32-byte box records)     -> alive[n]
//   2. A = inclusive prefix sum of alive over the preorder positions.  A subtree is a contiguous range [n, end[n]), so the
//      number of reachable primitives below the left / right operand of operator n is A[right-1] - A[n] / A[end-1] - A[right-1].
//   3. an operator survives when both sides have some; with one side empty it stands for the other side (Union; Difference whose
//      right side is empty) or is gone with everything below it (Difference without its left operand, Intersection with one
//      side: all M* cells, RaycastingKernels.cu:666-677).  The latter takes reachable primitives away from its ancestors:
//      clear them and repeat 2-3 (rare: needs an Intersection / a Difference that lost its left operand inside the tile).
//   4. S = inclusive prefix sum of the survivor flags = preorder numbering of the tile's tree: survivor n becomes record
//      S[n]-1, its left operand is the next record, its right operand record S[right[n]-1] (survivors before the right subtree).
//   5. primitives are emitted origin-relative (stage_record); operator boxes and the pure / bounded flags come from a bottom-up
//      refit of the tile's (small) tree in shared memory, with the same box rules as csg_prune_kernel (Union: both operands;
//      Difference: the left one; Intersection: the smaller).
// The trees differ from csg_prune_kernel's only where that one drops a whole operator by its box before looking at the
// primitives (never the other way round), and frames are byte-identical with either (tests/test_gpu_parity.py).
#ifdef CSG_PRUNE_PROBE   // one-off instrumented build (tools/gpu_prune_probe.py): globaltimer at the phase boundaries of every tile CTA
__device__ unsigned long long g_prune_probe[4096][16];
#define PROBE(k) do { if (threadIdx.x == 0 && blockIdx.x < 4096) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_prune_probe[blockIdx.x][k] = t_; } } while (0)
#else
#define PROBE(k) do { } while (0)
#endif
constexpr int kFlatThreadsMax = 512;   // CTA sizes: 128 (many tiles: throughput), 256, 512 (few tiles: the per-tile latency is the frame's fixed cost)
constexpr int kFlatMaxNodes = 32768;   // 2 x 16-bit prefix sums per node in shared memory (128 KB at the limit)
constexpr int kFlatPreferNodes = 2048;  // above this the tree-walking kernel is the faster one (it looks only at what the frustum touches)

struct FlatTileSmem {                  // followed by uint16_t A[n_pad], S[n_pad]
    float box[kSlotMax][6];            // culling box of every record of the tile's tree, origin-relative
    uint32_t meta[kSlotMax];           // kind | right operand (record index) << 8
    unsigned char flg[kSlotMax];       // bit0 pure, bit1 bounded, bit2 spheres and Unions only
    unsigned char cnt[kSlotMax];       // primitives below (saturating), for the flat flag
    uint32_t lmask[kSlotMax];          // bit j: record i + j of this subtree is a primitive (meaningful up to 32 records: 16 primitives)
    unsigned char done[kSlotMax];      // box and flags are final
    unsigned int wsum[kFlatThreadsMax / 32];
    float plane[5][4];                 // the tile's frustum
    unsigned int list_pos;             // this tile's entry in the bucket lists
};

// Inclusive prefix sum of v[0, n) in place by the whole CTA; every thread owns `chunk` consecutive elements (odd: the
// strided accesses then fall into distinct banks).  Ends with a barrier.
template <int T>
__device__ __forceinline__ void cta_scan_u16(unsigned short* v, int n, int chunk, unsigned int* wsum)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = min(tid * chunk, n), e = min(b + chunk, n);
    unsigned int s = 0;
    for (int i = b; i < e; ++i) s += v[i];
    unsigned int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned int run = inc - s;
#pragma unroll
    for (int k = 0; k < T / 32; ++k) run += k < warp ? wsum[k] : 0u;
    for (int i = b; i < e; ++i) { run += v[i]; v[i] = (unsigned short)run; }
    __syncthreads();
}

template <int T>
__global__ void __launch_bounds__(T, 1024 / T) csg_prune_flat_kernel(const __grid_constant__ PruneParams q)
{
    extern __shared__ __align__(16) unsigned char psm[];
    const int tid = threadIdx.x;
    const float ox = q.cam_pos[0], oy = q.cam_pos[1], oz = q.cam_pos[2];
    const int N = q.n_nodes, S = q.slot_nodes;
    cudaTriggerProgrammaticLaunchCompletion();   // the frame kernel may be scheduled now; it waits for this grid before it reads our output
    gate_enter(q.gate);
    if (blockIdx.x == 0 && tid < kCostBuckets && q.hist_next) q.hist_next[tid] = 0u;   // the bucket counters of the NEXT pruned frame (this frame's: q.hist)

    if (blockIdx.x < 2u) {
        // the first two CTAs ask the L2 for everything this kernel and the frame kernel read from the scene (leaf boxes, topology,
        // node records; primitive records): one request per 128-byte line, answered while the frustum is worked out
        auto pf = [tid](const void* base, size_t bytes, unsigned int part) {
            const char* pc = static_cast<const char*>(base);
            for (size_t o = ((size_t)part * T + tid) * 128u; o < bytes; o += 2u * T * 128u)
                asm volatile("prefetch.global.L2 [%0];" :: "l"(pc + o));
        };
        pf(q.leaf_boxes, (size_t)q.n_leaves * 32u, blockIdx.x);
        pf(q.topo, (size_t)N * 8u, blockIdx.x);
        pf(q.nodes, (size_t)N * 32u, blockIdx.x);
        if (q.prims) pf(q.prims, (size_t)q.n_prims * 80u, blockIdx.x);
    }
    if ((int)blockIdx.x >= q.n_tiles) {
        // staging CTAs: origin-relative copy of the whole tree at the head of the pool, for tiles whose tree does not fit a slot
        const int nb = (int)gridDim.x - q.n_tiles;
        for (int i = ((int)blockIdx.x - q.n_tiles) * T + tid; i < N; i += nb * T) {
            uint4 oa, ob;
            stage_record(__ldg(&q.nodes[2 * i]), __ldg(&q.nodes[2 * i + 1]), ox, oy, oz, oa, ob);
            q.pool[2 * i] = oa;
            q.pool[2 * i + 1] = ob;
        }
        return;
    }
    const int tile = (int)blockIdx.x;
    PROBE(0);
    FlatTileSmem& w = *reinterpret_cast<FlatTileSmem*>(psm);
    const int n_pad = (N + 7) & ~7;
    unsigned short* A = reinterpret_cast<unsigned short*>(psm + sizeof(FlatTileSmem));
    unsigned short* Sv = A + n_pad;

    int mx, my;
    tile_of(q, tile, mx, my);
    const int slot = slot_of_macro(q.shard_mode, mx, my, q.macro_x, q.shard_count);
    const int chunk = ((N + T - 1) / T) | 1;
    // the frustum is worked out by the first warp only (divisions, cross products: ~150 dependent instructions), while the
    // other warps clear the flags; everybody then keeps the five planes in registers, with the absolute values of the
    // normals: box [c - e, c + e] is outside plane n when n.c + |n|.e < 0
    if (tid < 32) {
        float pn0[5][3];
        tile_frustum(q, mx, my, pn0);
        if (tid < 5) { w.plane[tid][0] = pn0[tid][0]; w.plane[tid][1] = pn0[tid][1]; w.plane[tid][2] = pn0[tid][2]; }
    }
    for (int i = tid; i < N; i += T) A[i] = 0;
    __syncthreads();
    float pn[5][3], pa[5][3];
#pragma unroll
    for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) { pn[c][k] = w.plane[c][k]; pa[c][k] = fabsf(pn[c][k]); }
    PROBE(1);

    // ---- 1. reachable primitives
    for (int k0 = tid; k0 < q.n_leaves; k0 += 4 * T) {   // batches of four: all eight loads of a batch in flight together
        float4 la[4], lb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = min(k0 + j * T, q.n_leaves - 1);
            la[j] = __ldg(&q.leaf_boxes[2 * k]);
            lb[j] = __ldg(&q.leaf_boxes[2 * k + 1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // centre (relative to the camera) and half extent of the culling box
            const float ex = 0.5f * (lb[j].x - la[j].x), ey = 0.5f * (lb[j].y - la[j].y), ez = 0.5f * (lb[j].z - la[j].z);
            const float cx = (la[j].x - ox) + ex, cy = (la[j].y - oy) + ey, cz = (la[j].z - oz) + ez;
            bool outside = false;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const float m = fmaf(pn[c][0], cx, fmaf(pn[c][1], cy, pn[c][2] * cz)) + fmaf(pa[c][0], ex, fmaf(pa[c][1], ey, pa[c][2] * ez));
                outside = outside || (m < 0.0f);
            }
            if (k0 + j * T < q.n_leaves && !outside) A[__float_as_int(la[j].w)] = 1;
        }
    }
    __syncthreads();

    PROBE(2);
    // ---- 2./3. reachable primitives per subtree, survivors; primitives below an operator that is gone are taken away
    for (;;) {
        cta_scan_u16<T>(A, N, chunk, w.wsum);
        bool gone = false;
        for (int n0 = tid; n0 < N; n0 += 8 * T) {   // batches of eight: the topology loads of a batch are in flight together
            uint2 tp[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tp[j] = __ldg(&q.topo[min(n0 + j * T, N - 1)]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + j * T;
                if (n >= N) break;
                const uint32_t kind = tp[j].x & 7u;
                const unsigned int an = A[n];
                unsigned int surv;
                if (kind >= 3u) {
                    surv = an - (n ? (unsigned int)A[n - 1] : 0u);
                } else {
                    const unsigned int ar = A[(tp[j].x >> 8) - 1u], ae = A[tp[j].y - 1u];
                    const bool hl = ar != an, hr = ae != ar;   // something reachable below the left / right operand
                    surv = (hl && hr) ? 1u : 0u;
                    if (kind == 1u ? (!hl && hr) : (kind == 2u && hl != hr)) gone = true;
                }
                Sv[n] = (unsigned short)surv;
            }
        }
        if (!__syncthreads_or(gone ? 1 : 0)) break;
        // Sv holds the alive flag of every primitive: clear those below the operators that are gone, rebuild A from the flags.
        // Two parallel passes: the operators that are gone are marked (Sv = 2), then every primitive that is still alive walks up
        // its ancestors (a dozen cached loads) and drops out when one of them is marked.  (Until the end of round 2 the thread that
        // found an operator gone cleared its whole subtree by itself: 23 us for a tile that sees Cheese512's spheres but not its
        // cube — the Difference at the root goes, 1 024 nodes are cleared one by one.)
        for (int n = tid; n < N; n += T) {
            const uint2 tp = __ldg(&q.topo[n]);
            const uint32_t kind = tp.x & 7u;
            if (kind >= 3u) continue;
            const unsigned int an = A[n], ar = A[(tp.x >> 8) - 1u], ae = A[tp.y - 1u];
            const bool hl = ar != an, hr = ae != ar;
            if (kind == 1u ? (!hl && hr) : (kind == 2u && hl != hr)) Sv[n] = 2;
        }
        __syncthreads();
        for (int n = tid; n < N; n += T) {
            if ((__ldg(&q.topo[n]).x & 7u) < 3u || !Sv[n]) continue;
            for (int a = __ldg(&q.parent[n]); a >= 0; a = __ldg(&q.parent[a]))
                if (Sv[a] == 2) { Sv[n] = 0; break; }
        }
        __syncthreads();
        for (int n = tid; n < N; n += T) A[n] = ((__ldg(&q.topo[n]).x & 7u) >= 3u) ? Sv[n] : (unsigned short)0;
        __syncthreads();
    }

    PROBE(3);
    // ---- 4. preorder numbering of the survivors
    cta_scan_u16<T>(Sv, N, chunk, w.wsum);
    PROBE(4);
    const uint32_t kept = Sv[N - 1];
    const bool overflow = kept > (uint32_t)S;
    // heavy tiles (more nodes) are handed out first by the frame kernel: bucket lists now (the round trips of the atomic and
    // the list entry overlap the emission below), one ordered list at the end
    if (tid == 0 && q.lists) {
        const uint32_t cost = overflow ? (uint32_t)min(N, 2 * kCostBuckets - 1) : kept;
        const int bucket = (int)min(cost / 2u, (uint32_t)(kCostBuckets - 1));
        w.list_pos = (unsigned int)bucket * (unsigned int)q.n_slots + atomicAdd(&q.hist[bucket], 1u);   // filled in below
    }
    uint4* dst = q.pool + 2 * ((size_t)q.slots_off32 + (size_t)slot * S);
    uint32_t flags = 0u;
    if (!overflow && kept) {
        // ---- 5. records of the primitives, shape of the tile's tree
        for (int n = tid; n < N; n += T) {
            const uint32_t sn = Sv[n], sp = n ? (uint32_t)Sv[n - 1] : 0u;
            if (sn == sp) continue;
            const uint32_t i = sp;   // this survivor's record
            const uint2 tp = __ldg(&q.topo[n]);
            const uint32_t kind = tp.x & 7u;
            if (kind >= 3u) {
                const uint4 ua = __ldg(&q.nodes[2 * n]), ub = __ldg(&q.nodes[2 * n + 1]);
                uint4 oa, ob;
                stage_record(ua, ub, ox, oy, oz, oa, ob);
                dst[2 * i] = oa;
                dst[2 * i + 1] = ob;
                float lo[3], hi[3];
                rel_cull_box(ua, ub, ox, oy, oz, lo, hi);
#pragma unroll
                for (int c = 0; c < 3; ++c) { w.box[i][c] = lo[c]; w.box[i][3 + c] = hi[c]; }
                w.meta[i] = kind;
                w.flg[i] = (unsigned char)(((kind == 3u || kind == 5u) ? 1u : 0u) | ((kind != 4u) ? 2u : 0u) | ((kind == 3u) ? 4u : 0u));
                w.cnt[i] = 1;
                w.lmask[i] = 1u;
                w.done[i] = 1;
            } else {
                const uint32_t ri = Sv[(tp.x >> 8) - 1u];   // survivors before the right subtree = record of the right operand
                w.meta[i] = kind | (ri << 8);
                w.done[i] = 0;
            }
        }
        __syncthreads();
        PROBE(5);
        // ---- bottom-up refit in rounds, by the whole CTA: an operator whose operands are both done computes its box and flags,
        //      and is marked done behind a barrier (nobody reads a record that is being written); rounds = height of the tile's
        //      tree (a handful), shared memory only
        for (;;) {
            unsigned int now = 0u;
            bool pending = false;
            for (int i = tid, k = 0; i < (int)kept; i += T, ++k) {
                const uint32_t m = w.meta[i], kind = m & 7u;
                if (kind >= 3u || w.done[i]) continue;
                const int a = i + 1, b = (int)(m >> 8);
                if (!(w.done[a] && w.done[b])) { pending = true; continue; }
                const float* bl = w.box[a];
                const float* br = w.box[b];
                float* bo = w.box[i];
                if (kind == 0u) {                   // Union: both operands
#pragma unroll
                    for (int c = 0; c < 3; ++c) { bo[c] = fminf(bl[c], br[c]); bo[3 + c] = fmaxf(bl[3 + c], br[3 + c]); }
                } else if (kind == 1u) {            // Difference: a subset of the left operand
#pragma unroll
                    for (int c = 0; c < 6; ++c) bo[c] = bl[c];
                } else {                            // Intersection: a subset of both; the smaller box
                    float vl = 1.f, vr = 1.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) { vl *= fmaxf(bl[3 + c] - bl[c], 0.f); vr *= fmaxf(br[3 + c] - br[c], 0.f); }
                    const float* bs = vl <= vr ? bl : br;
#pragma unroll
                    for (int c = 0; c < 6; ++c) bo[c] = bs[c];
                }
                const uint32_t fl = w.flg[a], fr = w.flg[b];
                w.flg[i] = (unsigned char)(((kind == 0u) ? (fl & fr & 5u) : 0u) | (fl & fr & 2u));
                w.cnt[i] = (unsigned char)min((int)w.cnt[a] + (int)w.cnt[b], 255);
                w.lmask[i] = (w.lmask[a] << 1) | (w.lmask[b] << ((2u * w.cnt[a]) & 31u));   // the left subtree holds 2 cnt - 1 records, behind this one
                now |= 1u << k;
            }
            __syncthreads();
            for (int i = tid, k = 0; i < (int)kept; i += T, ++k)
                if ((now >> k) & 1u) w.done[i] = 1;
            if (!__syncthreads_or(pending ? 1 : 0)) break;   // (also the barrier in front of the operator records)
        }
        PROBE(6);
        // ---- records of the operators
        for (int i = tid; i < (int)kept; i += T) {
            const uint32_t m = w.meta[i], kind = m & 7u;
            if (kind >= 3u) continue;
            const uint32_t ri = m >> 8, f = w.flg[i];
            // flat: a Union over a few spheres (eval_flat_union, csg_scene.h); word 6 also tells Compute which operands are flat operators
            const int fmax = (int)kept <= q.flat_tree_max ? min(q.flat_max, kFlatLeavesMax) : 0;   // eval_flat_union reads the warp's shared-memory copy of the tree
            auto is_flat = [&w, fmax](uint32_t x) { return (w.meta[x] & 7u) == 0u && (w.flg[x] & 4u) && (int)w.cnt[x] <= fmax; };   // spheres-only Unions
            const bool flat = is_flat((uint32_t)i);
            const uint32_t meta = kind | (ri << 8) | ((w.meta[i + 1] & 7u) >= 3u ? kMetaLeftLeaf : 0u) | ((w.meta[ri] & 7u) >= 3u ? kMetaRightLeaf : 0u) |
                                  ((f & 2u) ? kMetaBounded : 0u) | ((f & 1u) ? kMetaPure : 0u) | (flat ? kMetaFlat : 0u);
            const uint32_t w6 = (flat ? ((w.lmask[i] >> 1) & kW6SphereMask) : 0u) | (is_flat((uint32_t)i + 1u) ? kW6LeftFlat : 0u) | (is_flat(ri) ? kW6RightFlat : 0u);
            const float* bo = w.box[i];
            dst[2 * i] = make_uint4(__float_as_uint(bo[0]), __float_as_uint(bo[1]), __float_as_uint(bo[2]), __float_as_uint(bo[3]));
            dst[2 * i + 1] = make_uint4(__float_as_uint(bo[4]), __float_as_uint(bo[5]), w6, meta);
        }
        const uint32_t rk = w.meta[0] & 7u;
        flags = (rk >= 3u ? kTileRootLeaf : 0u) | ((rk < 3u && (w.flg[0] & 1u)) ? kTileRootPure : 0u);
    }

    PROBE(7);
    // ---- descriptor and list entry (the entry carries the descriptor: the frame kernel gets the tile and its tree in one load)
    if (tid == 0) {
        const TileDesc td = overflow ? TileDesc{0u, (uint32_t)N, q.full_flags, 0u}
                                     : TileDesc{kept ? q.slots_off32 + (uint32_t)slot * (uint32_t)S : 0u, kept, flags, 0u};
        q.desc[slot] = td;
        if (q.lists) q.lists[w.list_pos] = make_uint4(td.offset32, td.n_nodes, td.flags, (uint32_t)tile);
    }
    PROBE(8);
    PROBE(9);
}

}  // namespace csgb
