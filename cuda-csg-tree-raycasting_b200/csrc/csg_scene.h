// Host-side scene model of libcsg_b200: parsed CSG tree + the flattened GPU layout.
//
// Replaces CSGTree / CSGNode / Primitive / BVHNode of the reference
// (RayCasting/CSGTree/CSGTree.cuh, CSGTree.cu, BVH/BVHNode.cuh, Primitives/Primitives.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace csgb {

enum NodeType : int { kUnion = 0, kDifference = 1, kIntersection = 2, kSphere = 3, kCylinder = 4, kCube = 5 };

// Same 44-byte / 48-byte layouts as the reference's CSGNode / Primitive so csg_scene_dump can hand
// them to a reference-side caller unchanged.
struct RefNode {
    int32_t type, prim, left, right, parent;
    float bmin[3], bmax[3];
};
struct RefPrim {
    int32_t id;
    float x, y, z;
    float r, g, b;
    float p[5];  // p[0] radius|size, p[1] height, p[2..4] axis
};
static_assert(sizeof(RefNode) == 44 && sizeof(RefPrim) == 48, "reference layouts");

// ---- flattened GPU tree -------------------------------------------------------------------
// One 32-byte record per node, preorder (left child of operator n is n+1):
//   word 7 (all kinds)   kind | leftIsLeaf<<3 | rightIsLeaf<<4 | bounded<<5 | pure<<6 | flat<<7 | idx<<8   idx = right child (operator) / primitive id (leaf)
//                        bounded = no cylinder below: the culling box really bounds every hit of the subtree (allows pruning)
//   operator   w0..5 = culling box (min xyz, max xyz),  w6 = parent node (-1 at the root) as uploaded; staged / tile trees: flat-evaluation word (kMetaFlat below)
//   sphere     w0..2 = centre, w3 = radius, w4..6 = centre again (staging turns w0..2 into origin-centre and w3 into r*r - |origin-centre|^2)
//   cube       w0..5 = lb, rt  (centre -/+ size/2, rounded like the reference rounds them)
//   cylinder   w0..5 = the reference's own leaf box (centre -/+ max(h/2, r)), which gates the primitive (Q6)
struct NodeRec {
    float f[7];
    uint32_t meta;
};
static_assert(sizeof(NodeRec) == 32, "record");
constexpr uint32_t kMetaLeftLeaf = 1u << 3, kMetaRightLeaf = 1u << 4, kMetaBounded = 1u << 5;
// pure = a Union whose subtree holds only Unions, spheres and cubes (every leaf convex and truly bounded by its box):
// such a subtree may be evaluated as a nearest-Enter search (csg_render.cu, ST_SEARCH)
constexpr uint32_t kMetaPure = 1u << 6;
// flat (per-tile trees only, set by csg_prune_flat_kernel) = a Union over at most PruneParams::flat_max (<= kFlatLeavesMax) spheres
// and nothing else: its result at any tmin follows from the spheres' roots alone (eval_flat_union, csg_kernel.cuh).  Word 6 of such a
// record, bits 0-28: the spheres among the records that follow it (bit j: record n + 1 + j; a subtree of k spheres is 2k - 1
// records).  Bits 30 / 31 of word 6 of ANY operator record of a tile tree say that its left / right operand is a flat operator
// (Compute looks there when it loops into an operand).  Records staged from the uploaded tree carry 0 in word 6.
// (Measured, round 2: flat Unions of up to 30 spheres — two flat operands scanned as one — are slower than two of 15 under an
// ordinary Union, whose box tests skip half the spheres: 0.170 against 0.166 ms per frame.)
constexpr uint32_t kMetaFlat = 1u << 7;
// cube and cylinder records of a staged / tile tree only (bit 3 means leftIsLeaf on operators): the six origin-relative bounds of
// the box are all within [2^-30, 2^30] in magnitude, so cube_isect / gate_box_exact may form their IEEE quotients from shared refined
// reciprocals (csg_kernel.cuh div_shared)
constexpr uint32_t kMetaBoxSafe = 1u << 3;
constexpr int kFlatLeavesMax = 15;
constexpr uint32_t kW6LeftFlat = 1u << 30, kW6RightFlat = 1u << 31, kW6SphereMask = (1u << 29) - 1u;

// Per-primitive data kept in global memory (read on accepted hits / cylinder + cube tests / shading): 5 x float4.
struct PrimRec {
    float color[4];   // r g b, kind
    float centre[4];  // x y z, p0 (radius | half size)
    float base[4];    // cylinder: C = centre - (h/2)V, height;  cube: |p - centre| thresholds of the normal components +-1, +-2 (0: none)
    float axis[4];    // cylinder: V, radius
    float haxis[4];   // cylinder: (h/2)V
};
static_assert(sizeof(PrimRec) == 80, "prim record");

struct FlatTree {
    std::vector<NodeRec> nodes;
    std::vector<PrimRec> prims;
    int depth = 0;        // operator levels on the longest path = stack slots the kernel needs
    bool root_is_leaf = false;
    bool root_pure = false;
    std::vector<int> parent;   // parent node of every record (-1 at the root), for the pruning kernel's upward marking
    std::vector<float> leaf_boxes;   // per primitive record, 8 floats: culling box min.xyz, node number (int bits), max.xyz, 0
    std::vector<uint32_t> subtree_end;   // per record: one past the last record of its subtree (preorder), for csg_prune_flat_kernel
    // Box outside which every ray is a Miss (the root's culling box; for a root primitive its true bounds, since a root leaf
    // is intersected without the reference's gating box, Q7).  Used for the per-frame screen-space bound.
    bool root_box_valid = false;
    float root_box[6] = {0, 0, 0, 0, 0, 0};
};

struct Scene {
    std::vector<RefNode> nodes;  // as parsed: the reference's tree, reference AABBs
    std::vector<RefPrim> prims;
    int optimize = 1;
    int depth() const;
};

// CSGTree::Parse.  Returns "" on success, else the reference's error text.
std::string parse_scene(const char* text, size_t len, Scene& out);
std::string write_scene(const Scene& s);
std::string generate_scene(int n_primitives, uint64_t seed);

// Smallest |p - centre| for which a component of the cube normal ((float)(int)((pc / half) * 1.00001f), RaycastingKernels.cu:422-424)
// reaches `level` (1 or 2); 0 when `half` is not a positive normal number.
float cube_normal_threshold_of(float half, float level);

// Builds the GPU layout.  optimize >= 1 re-balances Union-only subtrees spatially.
void flatten(const Scene& s, int optimize, FlatTree& out);

}  // namespace csgb
