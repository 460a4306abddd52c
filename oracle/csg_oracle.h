/* TEST INFRASTRUCTURE — not part of the product.
 *
 * csg_oracle: a plain-C, CPU restatement of the reference's per-pixel CSG raycast +
 * Phong path (Zumi002/CUDA-CSG-Tree-Raycasting, CSGRayCasting/Graphics/RayCasting/).
 * Every function in csg_oracle.c cites the reference file:line it follows.
 *
 * Arithmetic: the reference's golden is its CUDA kernel built by nvcc for sm_100 with
 * default flags (-fmad=true).  Which multiplies and adds fuse into FFMA is therefore part
 * of the observable behaviour (exact `t` ties decide pixels).  The
 * oracle is compiled with -ffp-contract=off and places fmaf() exactly where the
 * reference kernel's SASS has FFMA (decoded from `cuobjdump -sass` of the reference
 * build; see DESIGN.md §"Arithmetic contract").
 *
 * Pinning: tests/test_oracle_vs_reference.py checks this oracle against
 *   - oracle/_ref/libref_cpu.so  (the reference source itself, host build) in this container, and
 *   - tests/golden/ (.npz)        (outputs of oracle/_ref/libref_gpu.so = the reference CUDA
 *                                  kernel itself, generated on a B200 by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (libcsg_b200.so) never does.
 */
#ifndef CSG_ORACLE_H
#define CSG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CSGTree::NodeType, CSGTree.cuh:38-46 */
enum { ORC_UNION = 0, ORC_DIFFERENCE = 1, ORC_INTERSECTION = 2, ORC_SPHERE = 3, ORC_CYLINDER = 4, ORC_CUBE = 5 };

/* CSGNode, CSGTree.cuh:15-34 (44 bytes, same field order) */
typedef struct orc_node {
    int32_t type, prim, left, right, parent;
    float bmin[3], bmax[3]; /* BVHNode minX,minY,minZ,maxX,maxY,maxZ (BVHNode.cuh:10-15) */
} orc_node;

/* Primitive, Primitives.h:29-41 (48 bytes, same field order).
 * p[0] = radius (sphere, cylinder) or size (cube); p[1] = height; p[2..4] = axis (cylinder). */
typedef struct orc_prim {
    int32_t id;
    float x, y, z;
    float r, g, b;
    float p[5];
} orc_prim;

typedef struct orc_scene {
    int32_t n_nodes, n_prims;
    orc_node* nodes;
    orc_prim* prims;
} orc_scene;

/* Camera, Camera.h:9-16 (60 bytes, same field order) */
typedef struct orc_camera {
    float x, y, z;
    float rotX, rotY;
    float fov;
    float forward[3], right[3], up[3];
} orc_camera;

/* Per-ray event counts of the REFERENCE algorithm (the fixed yard-stick
 * the roofline's algorithmic flop/ray is computed from). */
typedef struct orc_counters {
    uint64_t rays, hits;
    uint64_t iters;      /* while-loop iterations, RaycastingKernels.cu:474 */
    uint64_t goto_calls; /* GoTo, :514 */
    uint64_t compute_calls; /* Compute, :597 */
    uint64_t aabb;       /* isBVHNodeHit, :718 */
    uint64_t sphere, cylinder, cube; /* sphereHit/cylinderHit/cubeHit */
    uint64_t max_iters;
    int32_t max_action_depth, max_hit_depth, max_time_depth;
    int32_t stack_overflows; /* pushes dropped by CudaStack (Q10) */
} orc_counters;

/* CSGTree::Parse, CSGTree.cu:5-152.  Returns 0 and *out on success; 1 and the
 * reference's exception text in err on failure. */
int orc_parse(const char* text, orc_scene** out, char* err, int errlen);
void orc_free(orc_scene* s);

/* Camera() + setPosition + setRotation (+ optional fov), Camera.h:18-35, Camera.cpp:4-36.
 * fov <= 0 keeps the reference default 90*3.14159/180. */
void orc_camera_init(orc_camera* cam, float x, float y, float z, float pitch, float yaw, float fov);
/* DirectionalLight::getLightDir, DirectionalLight.h:13-18; polar > 1e9 uses the defaults (:10-11). */
void orc_light_dir(float polar, float azimuth, float out3[3]);

/* Renders rows [y0,y1) (row 0 = bottom scanline) with the reference algorithm.
 *  hit[w*h]     RayHit.hit (0/1)                       RaycastingKernels.cu:36-46
 *  prim[w*h]    RayHit.primitiveIdx, -1 on miss
 *  t[w*h]       RayHit.t, -1 on miss
 *  flags[w*h]   RayHitMinimal.hit bit flags of the result (CSGUtils.cuh:29-37)
 *  rgba[w*h*4]  the float4 LightningKernel writes      RaycastingKernels.cu:49-111
 * Any output may be NULL.  tan_half_fov: pass NaN to compute tanf(fov/2) with the host libm, or the
 * value the CUDA device tanf returns to stay bit-identical with the GPU golden.
 * nthreads <= 0: all OpenMP threads.  counters may be NULL. */
int orc_render(const orc_scene* s, int w, int h, const orc_camera* cam, const float light_dir[3],
               float tan_half_fov, int y0, int y1, int nthreads,
               uint8_t* hit, int32_t* prim, float* t, uint8_t* flags, float* rgba,
               orc_counters* counters);

/* Single-primitive entry points for unit tests (same code the renderer uses).
 * out: t, flags(hit byte), id  — returns 1 if hit != Miss. */
int orc_hit_primitive(const orc_prim* p, int type, const float origin[3], const float dir[3], float tmin,
                      float* t_out, int* flags_out);
int orc_aabb_hit(const float bmin[3], const float bmax[3], const float origin[3], const float dir[3], float tmin);
/* Ray generation for one pixel (RaycastingKernels.cu:11-27 + Ray ctor, Ray.cuh:12-18). */
void orc_raygen(const orc_camera* cam, int w, int h, int x, int y, float tan_half_fov, float dir_out[3]);

#ifdef __cplusplus
}
#endif
#endif
