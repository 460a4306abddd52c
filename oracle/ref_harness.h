/* TEST INFRASTRUCTURE — not part of the product.
 *
 * C ABI shared by the two builds of the *unmodified reference* that live in
 * oracle/_ref/ (built by oracle/Makefile straight from /root/reference, never
 * copied into this repository):
 *
 *   libref_cpu.so  reference .cu sources compiled for the host with g++/OpenMP
 *                  (symbols refcpu_*)  -> CPU baseline + CPU-side pin of the oracle
 *   libref_gpu.so  reference .cu sources compiled with nvcc for sm_100
 *                  (symbols refgpu_*)  -> the primary golden (BASELINE.json north_star)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load these libraries.
 */
#ifndef REF_HARNESS_H
#define REF_HARNESS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One view of a scene.  Fields mirror the reference's parameter structs:
 * Camera{x,y,z,rotX,rotY,fov} (RenderManager/Camera/Camera.h:9-11) and
 * DirectionalLight{polar,azimuth} (RenderManager/DirectionalLight.h:10-11). */
typedef struct ref_view {
    int   width, height;
    float pos[3];
    float pitch, yaw;      /* radians, passed through Camera::setRotation (clamps pitch) */
    float fov;             /* radians; <= 0 keeps the reference default 90*3.14159/180 */
    float polar, azimuth;  /* radians; polar > 1e9 keeps the reference defaults */
} ref_view;

#define REF_DEFAULT_LIGHT 1e10f

#ifdef __cplusplus
}
#endif
#endif
