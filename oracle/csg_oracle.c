/* TEST INFRASTRUCTURE — not part of the product.  See csg_oracle.h.
 *
 * Plain-C restatement of the reference algorithm.  Citations are to
 * /root/reference/CSGRayCasting/Graphics/... ("RC/" = RayCasting/, "RM/" = RenderManager/).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (oracle/Makefile).  fmaf() appears exactly
 * where the reference kernel's sm_100 SASS has an FFMA; every other operation rounds
 * separately, like the FADD/FMUL/div.rn/sqrt.rn in that SASS.
 */
#define _GNU_SOURCE
#include "csg_oracle.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ enums */
/* CSG::CSGActions, RC/Utils/CSGUtils.cuh:5-13 */
enum { A_GOTO_LFT = 1 << 1, A_GOTO_RGH = 1 << 2, A_COMPUTE = 1 << 3, A_LOAD_LFT = 1 << 4, A_LOAD_RGH = 1 << 5, A_SAVE_LFT = 1 << 6 };
/* CSG::HitActions, CSGUtils.cuh:15-27 */
enum {
    H_MISS = 1 << 1, H_RETL = 1 << 2, H_RETR = 1 << 3, H_LOOPL = 1 << 4, H_LOOPR = 1 << 5,
    H_LOOPR_IF_CLOSER = 1 << 6, H_LOOPL_IF_CLOSER = 1 << 7, H_RETL_IF_CLOSER = 1 << 8, H_RETR_IF_CLOSER = 1 << 9, H_FLIPR = 1 << 10
};
/* CSG::CSGRayHit, CSGUtils.cuh:29-37 */
enum { R_ENTER = 1 << 1, R_EXIT = 1 << 2, R_MISS = 1 << 3, R_FLIP = 1 << 4, R_FLAG1 = 1 << 5, R_FLAG2 = 1 << 6 };

#define MAXSTACKSIZE 32 /* RC/Kernels/RaycastingKernels.cuh:19 */

typedef struct { float x, y, z; } f3;

/* RayHitMinimal, RC/Utils/Ray.cuh:36-52 */
typedef struct {
    float t;
    unsigned char hit;
    unsigned char primitiveType;
    short primitiveIdx;
} hitmin;

static hitmin hitmin_default(void)
{ /* Ray.cuh:45-50; primitiveType is left uninitialised by the reference, we use 0 */
    hitmin h;
    h.t = -1;
    h.hit = R_MISS;
    h.primitiveType = 0;
    h.primitiveIdx = -1;
    return h;
}

/* RayHit, Ray.cuh:24-34 */
typedef struct {
    int hit;
    float t;
    f3 position;
    f3 normal;
    int primitiveIdx;
} rayhit;

typedef struct { f3 origin, direction; } ray_t;

/* ------------------------------------------------------------------ stacks */
/* CudaStack<T,N>, RC/Utils/CudaStack.cuh:5-45: push on a full stack is dropped, pop on an
 * empty stack returns stack[0] without decrementing (Q10). */
#define DEF_STACK(NAME, T)                                                          \
    typedef struct { int count; T stack[MAXSTACKSIZE]; } NAME;                      \
    static int NAME##_push(NAME* s, T e)                                            \
    {                                                                               \
        if (s->count >= MAXSTACKSIZE) return 1;                                     \
        s->stack[s->count] = e;                                                     \
        s->count++;                                                                 \
        return 0;                                                                   \
    }                                                                               \
    static T NAME##_pop(NAME* s)                                                    \
    {                                                                               \
        if (s->count == 0) return s->stack[s->count];                               \
        s->count--;                                                                 \
        return s->stack[s->count];                                                  \
    }
DEF_STACK(stk_u8, unsigned char)
DEF_STACK(stk_hit, hitmin)
DEF_STACK(stk_f32, float)

/* ------------------------------------------------------------------ float helpers */
/* dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z (RC/Utils/Float3Utils.cuh:6-9) as compiled for
 * sm_100: FMUL(y) ; FFMA(x) ; FFMA(z). */
static inline float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
static inline f3 sub3(f3 a, f3 b) { f3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline f3 neg3(f3 a) { f3 r = {-a.x, -a.y, -a.z}; return r; }
/* normalize: invLen = 1.0f / sqrt(dot(a,a)); invLen * a  (Float3Utils.cuh:33-37; sqrt.rn + rcp.rn + FMUL) */
static inline f3 normalize3(f3 a)
{
    float inv = 1.0f / sqrtf(dot3(a, a));
    f3 r = {inv * a.x, inv * a.y, inv * a.z};
    return r;
}
/* Ray::computePosition: origin + t*direction (Ray.cuh:19-21) -> FFMA per component */
static inline f3 ray_at(const ray_t* r, float t)
{
    f3 p = {fmaf(t, r->direction.x, r->origin.x), fmaf(t, r->direction.y, r->origin.y), fmaf(t, r->direction.z, r->origin.z)};
    return p;
}
/* (float)(int)x as compiled: cvt.rzi.s32.f32 (saturating, NaN -> 0) then cvt.rn.f32.s32 */
static inline float trunc_via_int(float x)
{
    int i;
    if (x != x) i = 0;
    else if (x >= 2147483648.0f) i = 2147483647;
    else if (x <= -2147483648.0f) i = (-2147483647 - 1);
    else i = (int)x;
    return (float)i;
}

/* ------------------------------------------------------------------ parser */
/* color(hex), RC/CSGTree/CSGTree.cu:196-200: std::stoi(hex, nullptr, 16) then (float)value / 255 */
static int parse_color(const char* two, float* out, char* err, int errlen)
{
    char buf[3] = {two[0], two[1], 0};
    char* end = NULL;
    errno = 0;
    long v = strtol(buf, &end, 16);
    if (end == buf) { snprintf(err, errlen, "stoi"); return 1; }
    *out = (float)(int)v / 255;
    return 0;
}
/* std::stof / std::stod semantics: prefix parse, "stof"/"stod" on no conversion or range error */
static int parse_f(const char* tok, float* out, char* err, int errlen)
{
    char* end = NULL;
    errno = 0;
    float v = strtof(tok, &end);
    if (end == tok || errno == ERANGE) { snprintf(err, errlen, "stof"); return 1; }
    *out = v;
    return 0;
}
static int parse_d(const char* tok, double* out, char* err, int errlen)
{
    char* end = NULL;
    errno = 0;
    double v = strtod(tok, &end);
    if (end == tok || errno == ERANGE) { snprintf(err, errlen, "stod"); return 1; }
    *out = v;
    return 0;
}

/* BVHNode(Primitive, type), RC/CSGTree/BVH/BVHNode.cuh:17-66 */
static void leaf_box(const orc_prim* p, int type, float bmin[3], float bmax[3])
{
    float c[3] = {p->x, p->y, p->z};
    for (int i = 0; i < 3; ++i) bmin[i] = bmax[i] = 0;
    if (type == 3) { /* :19-33 sphere: centre -/+ r -/+ r */
        float r = p->p[0];
        for (int i = 0; i < 3; ++i) { bmin[i] = c[i] - r - r; bmax[i] = c[i] + r + r; }
    } else if (type == 4) { /* :34-49 cylinder: centre -/+ max(h/2, r), orientation ignored (Q6) */
        float half = p->p[1] / 2, rad = p->p[0];
        float maxR = half < rad ? rad : half; /* std::max(a,b) = (a<b)?b:a */
        for (int i = 0; i < 3; ++i) { bmin[i] = c[i] - maxR; bmax[i] = c[i] + maxR; }
    } else if (type == 5) { /* :50-65 cube */
        float halfSize = p->p[0] / 2;
        for (int i = 0; i < 3; ++i) { bmin[i] = c[i] - halfSize; bmax[i] = c[i] + halfSize; }
    }
}

/* CSGTree::ConstructBVH, CSGTree.cu:154-179 (post-order; operator box = union of child boxes
 * for every operator type, BVHNode.cuh:68-77) */
static void construct_bvh(orc_scene* s)
{
    /* preorder layout => children have larger indices than parents: a reverse sweep is post-order */
    for (int id = s->n_nodes - 1; id >= 0; --id) {
        orc_node* n = &s->nodes[id];
        if (n->prim != -1) {
            leaf_box(&s->prims[n->prim], n->type, n->bmin, n->bmax);
        } else {
            const orc_node* l = &s->nodes[n->left];
            const orc_node* r = &s->nodes[n->right];
            for (int i = 0; i < 3; ++i) {
                n->bmin[i] = r->bmin[i] < l->bmin[i] ? r->bmin[i] : l->bmin[i]; /* std::min(L,R) */
                n->bmax[i] = l->bmax[i] < r->bmax[i] ? r->bmax[i] : l->bmax[i]; /* std::max(L,R) */
            }
        }
    }
}

/* Primitive(...) cylinder constructor, RC/CSGTree/Primitives/Primitives.h:56-88 */
static void cylinder_axis(double rotX, double rotY, double rotZ, float axis[3])
{
    rotX = rotX * 0.017453292519943295769236907684886;
    rotY = rotY * 0.017453292519943295769236907684886;
    rotZ = rotZ * 0.017453292519943295769236907684886;
    double axisX = -sin(rotZ) * cos(rotY) + sin(rotY) * sin(rotX) * cos(rotZ),
           axisY = cos(rotX) * cos(rotZ),
           axisZ = sin(rotY) * sin(rotZ) + sin(rotX) * cos(rotY) * cos(rotZ);
    double len = (axisX * axisX + axisY * axisY + axisZ * axisZ); /* squared length, :80 */
    axisX /= len;
    axisY /= len;
    axisZ /= len;
    axis[0] = (float)axisX;
    axis[1] = (float)axisY;
    axis[2] = (float)axisZ;
}

/* CSGTree::Parse, CSGTree.cu:5-152 (+ split :181-194) */
int orc_parse(const char* text, orc_scene** out, char* err, int errlen)
{
    char dummy[8];
    if (!err || errlen <= 0) { err = dummy; errlen = (int)sizeof dummy; }
    err[0] = 0;
    *out = NULL;
    /* split(): whitespace-separated tokens */
    size_t len = strlen(text);
    char* buf = (char*)malloc(len + 1);
    memcpy(buf, text, len + 1);
    int ntok = 0, cap = 64;
    char** tok = (char**)malloc(sizeof(char*) * cap);
    for (size_t i = 0; i < len;) {
        while (i < len && isspace((unsigned char)buf[i])) buf[i++] = 0;
        if (i >= len) break;
        if (ntok == cap) { cap *= 2; tok = (char**)realloc(tok, sizeof(char*) * cap); }
        tok[ntok++] = buf + i;
        while (i < len && !isspace((unsigned char)buf[i])) ++i;
    }

    orc_scene* s = (orc_scene*)calloc(1, sizeof *s);
    s->nodes = (orc_node*)calloc((size_t)ntok + 1, sizeof(orc_node));
    s->prims = (orc_prim*)calloc((size_t)ntok + 1, sizeof(orc_prim));
    /* nodesStack: (nodeIdx, childrenCount) pairs, :11 */
    int* st_node = (int*)malloc(sizeof(int) * ((size_t)ntok + 1));
    int* st_cnt = (int*)malloc(sizeof(int) * ((size_t)ntok + 1));
    int sp = 0;
    int primitivesCount = 0, nodesCount = 0;
    int rc = 0;

#define FAIL(...) do { snprintf(err, errlen, __VA_ARGS__); rc = 1; goto done; } while (0)
#define NEED(k) do { if (i + (k) >= ntok) FAIL("Cannot parse - unexpected end of input"); } while (0) /* reference reads out of range (UB) */

    for (int i = 0; i < ntok; i++) {
        orc_node* nd = &s->nodes[nodesCount];
        nd->type = nd->prim = nd->left = nd->right = nd->parent = -1; /* :19 */
        if (nodesCount != 0) { /* :20-40 */
            if (sp == 0) FAIL("Cannot parse");
            if (st_cnt[sp - 1] == 0) {
                s->nodes[st_node[sp - 1]].left = nodesCount;
                nd->parent = st_node[sp - 1];
                st_cnt[sp - 1]++;
            } else {
                s->nodes[st_node[sp - 1]].right = nodesCount;
                nd->parent = st_node[sp - 1];
                sp--;
            }
        }
        const char* kw = tok[i];
        if (!strcmp(kw, "Union") || !strcmp(kw, "Difference") || !strcmp(kw, "Intersection")) { /* :42-56 */
            nd->type = !strcmp(kw, "Union") ? ORC_UNION : !strcmp(kw, "Difference") ? ORC_DIFFERENCE : ORC_INTERSECTION;
            st_node[sp] = nodesCount;
            st_cnt[sp] = 0;
            sp++;
        } else if (!strcmp(kw, "Sphere") || !strcmp(kw, "Cube") || !strcmp(kw, "Cylinder")) { /* :57-137 */
            int is_cyl = !strcmp(kw, "Cylinder");
            nd->type = !strcmp(kw, "Sphere") ? ORC_SPHERE : is_cyl ? ORC_CYLINDER : ORC_CUBE;
            nd->prim = primitivesCount;
            orc_prim* p = &s->prims[primitivesCount];
            NEED(is_cyl ? 9 : 5);
            p->id = primitivesCount;
            if (parse_f(tok[i + 1], &p->x, err, errlen) || parse_f(tok[i + 2], &p->y, err, errlen) ||
                parse_f(tok[i + 3], &p->z, err, errlen)) { rc = 1; goto done; }
            if (strlen(tok[i + 4]) != 6) FAIL("Cannot parse color %s", tok[i + 4]);
            if (parse_color(tok[i + 4], &p->r, err, errlen) || parse_color(tok[i + 4] + 2, &p->g, err, errlen) ||
                parse_color(tok[i + 4] + 4, &p->b, err, errlen)) { rc = 1; goto done; }
            if (parse_f(tok[i + 5], &p->p[0], err, errlen)) { rc = 1; goto done; } /* radius / size */
            if (is_cyl) {
                double rx, ry, rz;
                if (parse_f(tok[i + 6], &p->p[1], err, errlen) || parse_d(tok[i + 7], &rx, err, errlen) ||
                    parse_d(tok[i + 8], &ry, err, errlen) || parse_d(tok[i + 9], &rz, err, errlen)) { rc = 1; goto done; }
                if (rx > 360 || rx < 0) FAIL("Invalid roation rotX should be in range [0, 360] deg"); /* :103-108, sic */
                if (ry > 360 || ry < 0) FAIL("Invalid roation rotY should be in range [0, 360] deg");
                if (rz > 360 || rz < 0) FAIL("Invalid roation rotZ should be in range [0, 360] deg");
                cylinder_axis(rx, ry, rz, &p->p[2]);
                i += 9;
            } else {
                i += 5;
            }
            primitivesCount++;
        } else {
            FAIL("Cannot parse - Unrecognized keyword: %s", kw); /* :138-141 */
        }
        nodesCount++;
    }
    if (nodesCount != 2 * primitivesCount - 1) /* :146-147 */
        FAIL("Cannot parse - number of primitives do not match number of nodes");
    s->n_nodes = nodesCount;
    s->n_prims = primitivesCount;
    construct_bvh(s);
done:
    free(st_node);
    free(st_cnt);
    free(tok);
    free(buf);
    if (rc) { orc_free(s); return rc; }
    *out = s;
    return 0;
#undef FAIL
#undef NEED
}

void orc_free(orc_scene* s)
{
    if (!s) return;
    free(s->nodes);
    free(s->prims);
    free(s);
}

/* ------------------------------------------------------------------ camera / light */
/* Camera::normalizeVector, RM/Camera/Camera.cpp:29-36: float sum, double sqrt (only <cmath>'s
 * ::sqrt(double) is visible), result stored to float, float divides */
static void cam_normalize(float* vec)
{
    float length = (float)sqrt((double)(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]));
    if (length > 0) {
        vec[0] /= length;
        vec[1] /= length;
        vec[2] /= length;
    }
}
/* Camera::updateVectors, Camera.cpp:4-27.  sin/cos resolve to the double overloads there. */
static void cam_update(orc_camera* c)
{
    c->forward[0] = (float)(-sin((double)c->rotY) * cos((double)c->rotX));
    c->forward[1] = (float)sin((double)c->rotX);
    c->forward[2] = (float)(-cos((double)c->rotY) * cos((double)c->rotX));
    cam_normalize(c->forward);
    c->right[0] = (float)cos((double)c->rotY);
    c->right[1] = 0;
    c->right[2] = (float)-sin((double)c->rotY);
    cam_normalize(c->right);
    c->up[0] = c->forward[1] * c->right[2] - c->forward[2] * c->right[1];
    c->up[1] = c->forward[2] * c->right[0] - c->forward[0] * c->right[2];
    c->up[2] = c->forward[0] * c->right[1] - c->forward[1] * c->right[0];
    cam_normalize(c->up);
}
void orc_camera_init(orc_camera* cam, float x, float y, float z, float pitch, float yaw, float fov)
{
    /* Camera(), Camera.h:18-20 */
    cam->fov = 90.0f * 3.14159f / 180.0f;
    /* setPosition, :22-26 */
    cam->x = x; cam->y = y; cam->z = z;
    /* setRotation, :28-35: clamp pitch to +-89 deg */
    cam->rotX = fmaxf(-89.0f * 3.14159f / 180.0f, fminf(89.0f * 3.14159f / 180.0f, pitch));
    cam->rotY = yaw;
    cam_update(cam);
    if (fov > 0) cam->fov = fov;
}
/* DirectionalLight::getLightDir, RM/DirectionalLight.h:8-18 (float sin/cos overloads) */
void orc_light_dir(float polar, float azimuth, float out3[3])
{
    if (polar > 1e9f) {
        polar = -60.f * 3.14159f / 180.f;
        azimuth = -45.f * 3.14159f / 180.f;
    }
    out3[0] = sinf(polar) * cosf(azimuth);
    out3[1] = cosf(polar);
    out3[2] = sinf(polar) * sinf(azimuth);
}

/* ------------------------------------------------------------------ ray generation */
/* RaycastKernel, RC/Kernels/RaycastingKernels.cu:11-27, and Ray::Ray, RC/Utils/Ray.cuh:12-18 */
static ray_t raygen(const orc_camera* cam, float width, float height, int x, int y, float th)
{
    float u = ((float)x + 0.5f) / (width - 1);           /* :11  div.rn */
    float v = ((float)y + 0.5f) / (height - 1);          /* :12 */
    float nx = ((width / height) * fmaf(u, 2.0f, -1.0f)) * th; /* :15  (2u-1) is one FFMA */
    float ny = (1.0f - (v + v)) * th;                    /* :16 */
    f3 c;
    c.x = cam->forward[0] + fmaf(cam->right[0], nx, cam->up[0] * ny); /* :22-24 */
    c.y = cam->forward[1] + fmaf(cam->right[1], nx, cam->up[1] * ny);
    c.z = cam->forward[2] + fmaf(cam->right[2], nx, cam->up[2] * ny);
    ray_t r;
    r.origin.x = cam->x; r.origin.y = cam->y; r.origin.z = cam->z; /* :20 */
    r.direction = normalize3(normalize3(c)); /* :21 normalize, then Ray ctor normalises again (Q3) */
    return r;
}
void orc_raygen(const orc_camera* cam, int w, int h, int x, int y, float tan_half_fov, float dir_out[3])
{
    float th = tan_half_fov == tan_half_fov ? tan_half_fov : tanf(cam->fov * 0.5f);
    ray_t r = raygen(cam, (float)w, (float)h, x, y, th);
    dir_out[0] = r.direction.x; dir_out[1] = r.direction.y; dir_out[2] = r.direction.z;
}

/* ------------------------------------------------------------------ primitives */
typedef struct {
    uint64_t iters, goto_calls, compute_calls, aabb, sphere, cylinder, cube;
    int max_a, max_p, max_t, overflow;
} evt;

/* sphereHit, RaycastingKernels.cu:135-181 */
static void sphere_hit(const ray_t* ray, const orc_prim* sp, hitmin* h, float tmin, evt* ev)
{
    ev->sphere++;
    h->hit = R_MISS;                     /* :137 */
    h->primitiveIdx = (short)sp->id;     /* :138 (int -> short, Q11) */
    f3 c = {sp->x, sp->y, sp->z};
    f3 oc = sub3(ray->origin, c);        /* :139-143 */
    float b = dot3(oc, ray->direction);  /* :145 */
    float radius = sp->p[0];
    /* :146-147  c = dot(oc,oc) - r*r ; disc = b*b - c.  SASS: FFMA(r,r,-dot) ; FFMA(b,b,.) */
    float negc = fmaf(radius, radius, -dot3(oc, oc));
    float discriminant = fmaf(b, b, negc);
    if (discriminant < 0) return;        /* :149 */
    float sq = sqrtf(discriminant);
    float temp = (-b) - sq;              /* :151 */
    if (temp <= tmin) {                  /* :152  (a NaN root is NOT rejected: the reference's own comparison) */
        temp = sq - b;                   /* :153 */
        if (temp <= tmin) {              /* :154-160 */
            h->t = -1;
            h->hit = R_MISS;
            h->primitiveIdx = -1;
            return;
        }
    }
    h->t = temp;                         /* :163 */
    f3 n = sub3(ray_at(ray, temp), c);   /* :165-171 */
    h->hit = (dot3(n, ray->direction) <= 0) ? R_ENTER : R_EXIT; /* :173-176 (NaN -> Exit) */
    h->primitiveType = ORC_SPHERE;       /* :178 */
}

/* cubeHit, RaycastingKernels.cu:375-434 */
static void cube_hit(const ray_t* ray, const orc_prim* cu, hitmin* h, float tmin, evt* ev)
{
    ev->cube++;
    h->hit = R_MISS;                     /* :377 */
    h->primitiveIdx = (short)cu->id;     /* :378 */
    f3 C = {cu->x, cu->y, cu->z};
    float hs = cu->p[0] / 2;             /* size/2 * axis(1,1,1): exact scaling */
    f3 lb = {C.x - hs, C.y - hs, C.z - hs}; /* :387 */
    f3 rt = {C.x + hs, C.y + hs, C.z + hs}; /* :388 */
    float t1 = (lb.x - ray->origin.x) / ray->direction.x; /* :389-394  div.rn */
    float t2 = (rt.x - ray->origin.x) / ray->direction.x;
    float t3 = (lb.y - ray->origin.y) / ray->direction.y;
    float t4 = (rt.y - ray->origin.y) / ray->direction.y;
    float t5 = (lb.z - ray->origin.z) / ray->direction.z;
    float t6 = (rt.z - ray->origin.z) / ray->direction.z;
    float tempmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6)); /* :396 */
    float tempmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6)); /* :397 */
    if (tempmax < 0) return;             /* :401 */
    if (tempmin > tempmax) return;       /* :407 */
    if (tempmin <= tmin) {               /* :411 */
        tempmin = tempmax;
        if (tempmin <= tmin) return;     /* :414 */
    }
    h->t = tempmin;                      /* :420 */
    f3 PC = sub3(ray_at(ray, tempmin), C); /* :421 */
    const float bias = 1.00001f;         /* :422 */
    f3 n = {trunc_via_int((PC.x / hs) * bias), trunc_via_int((PC.y / hs) * bias), trunc_via_int((PC.z / hs) * bias)}; /* :424 */
    /* :427  dot(normal, dir) — operand order irrelevant for FMUL/FFMA */
    h->hit = (dot3(n, ray->direction) <= 0) ? R_ENTER : R_EXIT;
    h->primitiveType = ORC_CUBE;         /* :432 */
}

/* cylinderHit, RaycastingKernels.cu:202-336.  FFMA placement follows the sm_100 SASS of the
 * reference build; in particular dot(d,V) here is FFMA(Vy,dy,Vx*dx) + Vz*dz
 * (the two products are shared with dot(d,-V)), while the other dots use the FMUL,FFMA,FFMA form. */
static void cylinder_hit(const ray_t* ray, const orc_prim* cy, hitmin* h, float tmin, evt* ev)
{
    ev->cylinder++;
    h->hit = R_MISS;                     /* :204 */
    h->primitiveIdx = (short)cy->id;     /* :205 */
    const float radius = cy->p[0], height = cy->p[1];
    const f3 V = {cy->p[2], cy->p[3], cy->p[4]}; /* :207 */
    const f3 d = ray->direction, o = ray->origin;
    const float hh = height * 0.5f;      /* height / 2 */
    const f3 hV = {V.x * hh, V.y * hh, V.z * hh};
    const f3 C = {cy->x - hV.x, cy->y - hV.y, cy->z - hV.z}; /* :208 */
    const f3 OC = sub3(o, C);            /* :209 */

    const float pxd = V.x * d.x, pzd = V.z * d.z;
    const float dV = fmaf(V.y, d.y, pxd) + pzd;           /* :211 aHelp */
    const float a = fmaxf(fmaf(-dV, dV, 1.0f), 0.00001f); /* :212 */
    const float OCV = fmaf(V.z, OC.z, fmaf(V.x, OC.x, V.y * OC.y)); /* :214 cHelp */
    const float OC2 = fmaf(OC.z, OC.z, fmaf(OC.x, OC.x, OC.y * OC.y));
    const float c = fmaf(-radius, radius, fmaf(-OCV, OCV, OC2)); /* :215 */
    const float dOC = fmaf(OC.z, d.z, fmaf(OC.x, d.x, OC.y * d.y));
    const float b = fmaf(dV, -OCV, dOC);                   /* :217 */
    const float discriminant = fmaf(b, b, -(a * c));      /* :218 */
    if (discriminant < 0) return;                          /* :220 */

    const float sq = sqrtf(discriminant);
    const float t1 = ((-b) - sq) / a, t2 = (sq - b) / a;   /* :224 */
    const float m1 = fmaf(dV, t1, OCV), m2 = fmaf(dV, t2, OCV); /* :225 */
    if ((m1 < 0 && m2 < 0) || (m1 > height && m2 > height)) return; /* :229 */

    /* den = dot(d,-V) (:240,:275): FFMA(-Vy,dy,-(Vx*dx)) - Vz*dz ;  dot(-OC,-V) == OCV exactly */
    const float den_bottom = fmaf(-V.y, d.y, -pxd) - pzd;
    /* OCmMax = origin - centre - (h/2)V ; dot(-OCmMax, V)  (:261-262, :296-297) */
    const float X = (o.x - cy->x) - hV.x, Y = (o.y - cy->y) - hV.y, Z = (o.z - cy->z) - hV.z;

    float temp = t1, m = m1;             /* :232-233 */
    int skip = 0, useSurfNormal = 0;     /* :235-237 */
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) { temp = t2; m = m2; skip = 0; useSurfNormal = 0; } /* :269-272 */
        if (m < 0) {                     /* :238-251 / :273-286 bottom cap */
            if (fabsf(den_bottom) < 0.0001f) skip = 1;
            else { temp = OCV / den_bottom; useSurfNormal = 1; }
        }
        if (m > height) {                /* :252-266 / :287-301 top cap */
            if (fabsf(dV) < 0.0001f) skip = 1;
            else { temp = fmaf(-V.z, Z, fmaf(V.y, -Y, -(V.x * X))) / dV; useSurfNormal = 2; }
        }
        if (!(temp <= tmin || skip)) break; /* :267 / :302 */
        if (pass == 1) return;           /* :304 */
    }

    h->t = temp;                         /* :310 */
    f3 normal;
    if (useSurfNormal == 1) normal = neg3(V);
    else if (useSurfNormal == 2) normal = V;
    else {                               /* :319  P - C - m*V  -> FFMA(-V, m, P - C) */
        f3 P = ray_at(ray, temp);
        normal.x = fmaf(-V.x, m, P.x - C.x);
        normal.y = fmaf(-V.y, m, P.y - C.y);
        normal.z = fmaf(-V.z, m, P.z - C.z);
    }
    /* :322  dot(normal, dir): FMUL(dy,ny) ; FFMA(dx,nx) ; FFMA(dz,nz) */
    h->hit = (dot3(normal, d) <= 0) ? R_ENTER : R_EXIT;
    if (useSurfNormal == 1) h->hit |= R_FLAG1;      /* :327-334 */
    else if (useSurfNormal == 2) h->hit |= R_FLAG2;
    h->primitiveType = ORC_CYLINDER;     /* :335 */
}

/* hitPrimitive, RaycastingKernels.cu:113-133 */
static void hit_primitive(const ray_t* ray, const orc_scene* s, const orc_node* node, hitmin* h, float tmin, evt* ev)
{
    if (node->type == ORC_SPHERE) sphere_hit(ray, &s->prims[node->prim], h, tmin, ev);
    else if (node->type == ORC_CYLINDER) cylinder_hit(ray, &s->prims[node->prim], h, tmin, ev);
    else if (node->type == ORC_CUBE) cube_hit(ray, &s->prims[node->prim], h, tmin, ev);
    else *h = hitmin_default();
}

/* isBVHNodeHit, RaycastingKernels.cu:718-757 */
static int bvh_hit(const ray_t* ray, const float bmin[3], const float bmax[3], hitmin* h, float tmin, evt* ev)
{
    ev->aabb++;
    float t1 = (bmin[0] - ray->origin.x) / ray->direction.x; /* :724-729 */
    float t2 = (bmax[0] - ray->origin.x) / ray->direction.x;
    float t3 = (bmin[1] - ray->origin.y) / ray->direction.y;
    float t4 = (bmax[1] - ray->origin.y) / ray->direction.y;
    float t5 = (bmin[2] - ray->origin.z) / ray->direction.z;
    float t6 = (bmax[2] - ray->origin.z) / ray->direction.z;
    float tempmin = fmaxf(fmaxf(fminf(t1, t2), fminf(t3, t4)), fminf(t5, t6)); /* :731 */
    float tempmax = fminf(fminf(fmaxf(t1, t2), fmaxf(t3, t4)), fmaxf(t5, t6)); /* :732 */
    if (tempmax < 0) { if (h) h->hit = R_MISS; return 0; }        /* :735-739 */
    if (tempmin > tempmax) { if (h) h->hit = R_MISS; return 0; }  /* :742-746 */
    if (tempmin <= tmin) {                                        /* :747-754 */
        if (tempmax <= tmin) { if (h) h->hit = R_MISS; return 0; }
    }
    return 1;
}

/* ------------------------------------------------------------------ state machine */
/* LookUpActions, RaycastingKernels.cu:664-706 */
static int lookup_actions(unsigned char lHit, unsigned char rHit, int op)
{
    static const int unionTable[3][3] = {
        {H_RETL_IF_CLOSER | H_RETR_IF_CLOSER, H_RETR_IF_CLOSER | H_LOOPL, H_RETL},
        {H_RETL_IF_CLOSER | H_LOOPR, H_LOOPL_IF_CLOSER | H_LOOPR_IF_CLOSER, H_RETL},
        {H_RETR, H_RETR, H_MISS}};
    static const int intersectionTable[3][3] = {
        {H_LOOPL_IF_CLOSER | H_LOOPR_IF_CLOSER, H_RETL_IF_CLOSER | H_LOOPR, H_MISS},
        {H_RETR_IF_CLOSER | H_LOOPL, H_RETL_IF_CLOSER | H_RETR_IF_CLOSER, H_MISS},
        {H_MISS, H_MISS, H_MISS}};
    static const int differenceTable[3][3] = {
        {H_RETL_IF_CLOSER | H_LOOPR, H_LOOPL_IF_CLOSER | H_LOOPR_IF_CLOSER, H_RETL},
        {H_RETL_IF_CLOSER | H_RETR_IF_CLOSER | H_FLIPR, H_RETR_IF_CLOSER | H_FLIPR | H_LOOPL, H_RETL},
        {H_MISS, H_MISS, H_MISS}};
    if (lHit & R_ENTER) lHit = 0; /* :679-684 (sequential ifs on the running value) */
    if (lHit & R_EXIT) lHit = 1;
    if (lHit & R_MISS) lHit = 2;
    if (rHit & R_ENTER) rHit = 0;
    if (rHit & R_EXIT) rHit = 1;
    if (rHit & R_MISS) rHit = 2;
    if (lHit > 2 || rHit > 2) return -1; /* reference would index out of bounds; never happens for valid hits */
    if (op == ORC_UNION) return unionTable[lHit][rHit];
    if (op == ORC_INTERSECTION) return intersectionTable[lHit][rHit];
    if (op == ORC_DIFFERENCE) return differenceTable[lHit][rHit];
    return -1;
}

static const orc_node VIRTUAL_NODE = {0, 0, 0, 0, 0, {0, 0, 0}, {0, 0, 0}}; /* CSGNode(0,0,0,0,0), :467, :715 */

/* GetParent, RaycastingKernels.cu:708-716 */
static orc_node get_parent(const orc_scene* s, const orc_node* node, int* run)
{
    if (node->parent >= 0) return s->nodes[node->parent];
    *run = 0;
    return VIRTUAL_NODE;
}

typedef struct {
    stk_u8 A;
    stk_hit P;
    stk_f32 T;
} stacks;

#define TRACK(ev, st) do { if ((st)->A.count > (ev)->max_a) (ev)->max_a = (st)->A.count; \
                           if ((st)->P.count > (ev)->max_p) (ev)->max_p = (st)->P.count; \
                           if ((st)->T.count > (ev)->max_t) (ev)->max_t = (st)->T.count; } while (0)

/* GoTo, RaycastingKernels.cu:514-595 */
static void go_to(stacks* st, unsigned char* action, orc_node* node, const orc_scene* s, hitmin* leftRay, hitmin* rightRay,
                  const ray_t* ray, float* tmin, int* run, evt* ev)
{
    ev->goto_calls++;
    if (*action & A_GOTO_LFT) *node = s->nodes[node->left];  /* :527-534 */
    else *node = s->nodes[node->right];

    if (node->type == ORC_UNION || node->type == ORC_DIFFERENCE || node->type == ORC_INTERSECTION) { /* :536 */
        const orc_node* ln = &s->nodes[node->left];
        const orc_node* rn = &s->nodes[node->right];
        int gotoL = bvh_hit(ray, ln->bmin, ln->bmax, leftRay, *tmin, ev);  /* :540 */
        int gotoR = bvh_hit(ray, rn->bmin, rn->bmax, rightRay, *tmin, ev); /* :541 */
        if (gotoL && ln->prim != -1) { hit_primitive(ray, s, ln, leftRay, *tmin, ev); gotoL = 0; }  /* :542-547 */
        if (gotoR && rn->prim != -1) { hit_primitive(ray, s, rn, rightRay, *tmin, ev); gotoR = 0; } /* :548-553 */
        if (gotoL || gotoR) {
            if (!gotoL) {            /* :556-561 */
                ev->overflow += stk_hit_push(&st->P, *leftRay);
                ev->overflow += stk_u8_push(&st->A, A_LOAD_LFT);
                *action = A_GOTO_RGH;
            } else if (!gotoR) {     /* :562-567 */
                ev->overflow += stk_hit_push(&st->P, *rightRay);
                ev->overflow += stk_u8_push(&st->A, A_LOAD_RGH);
                *action = A_GOTO_LFT;
            } else {                 /* :568-574 */
                ev->overflow += stk_f32_push(&st->T, *tmin);
                ev->overflow += stk_u8_push(&st->A, A_LOAD_LFT);
                ev->overflow += stk_u8_push(&st->A, A_SAVE_LFT);
                *action = A_GOTO_LFT;
            }
            TRACK(ev, st);
        } else {
            *action = A_COMPUTE;     /* :578 */
        }
    } else {                         /* :582-594 leaf reached directly: no box test (Q7) */
        if (*action & A_GOTO_LFT) hit_primitive(ray, s, node, leftRay, *tmin, ev);
        else hit_primitive(ray, s, node, rightRay, *tmin, ev);
        *action = stk_u8_pop(&st->A);
        *node = get_parent(s, node, run);
    }
}

/* Compute, RaycastingKernels.cu:597-661 */
static void compute(stacks* st, unsigned char* action, orc_node* node, const orc_scene* s, hitmin* leftRay, hitmin* rightRay,
                    float* tmin, int* run, evt* ev)
{
    ev->compute_calls++;
    if (*action & (A_LOAD_LFT | A_LOAD_RGH)) { /* :609-619 */
        if (*action & A_LOAD_LFT) *leftRay = stk_hit_pop(&st->P);
        else *rightRay = stk_hit_pop(&st->P);
    }
    int actions = lookup_actions(leftRay->hit, rightRay->hit, node->type); /* :620 */
    if ((actions & H_RETL) || ((actions & H_RETL_IF_CLOSER) && (leftRay->t < rightRay->t))) { /* :621-626 */
        *rightRay = *leftRay;
        *action = stk_u8_pop(&st->A);
        *node = get_parent(s, node, run);
    } else if ((actions & H_RETR) || ((actions & H_RETR_IF_CLOSER) && (leftRay->t > rightRay->t))) { /* :627-639 */
        if (actions & H_FLIPR) {
            rightRay->hit ^= R_FLIP;
            rightRay->hit ^= R_EXIT;
            rightRay->hit ^= R_ENTER;
        }
        *leftRay = *rightRay;
        *action = stk_u8_pop(&st->A);
        *node = get_parent(s, node, run);
    } else if ((actions & H_LOOPL) || ((actions & H_LOOPL_IF_CLOSER) && (leftRay->t < rightRay->t))) { /* :640-646 */
        *tmin = leftRay->t;
        ev->overflow += stk_hit_push(&st->P, *rightRay);
        ev->overflow += stk_u8_push(&st->A, A_LOAD_RGH);
        *action = A_GOTO_LFT;
        TRACK(ev, st);
    } else if ((actions & H_LOOPR) || ((actions & H_LOOPR_IF_CLOSER) && (leftRay->t > rightRay->t))) { /* :647-653 */
        *tmin = rightRay->t;
        ev->overflow += stk_hit_push(&st->P, *leftRay);
        ev->overflow += stk_u8_push(&st->A, A_LOAD_LFT);
        *action = A_GOTO_RGH;
        TRACK(ev, st);
    } else {                         /* :654-660 */
        *rightRay = hitmin_default();
        *leftRay = hitmin_default();
        *action = stk_u8_pop(&st->A);
        *node = get_parent(s, node, run);
    }
}

/* CSGRayCast, RaycastingKernels.cu:459-512 */
static hitmin csg_ray_cast(const orc_scene* s, const ray_t* ray, evt* ev)
{
    stacks st;
    st.A.count = st.P.count = st.T.count = 0;
    memset(st.A.stack, 0, sizeof st.A.stack); /* the reference leaves these uninitialised */
    memset(st.T.stack, 0, sizeof st.T.stack);
    for (int i = 0; i < MAXSTACKSIZE; ++i) st.P.stack[i] = hitmin_default();
    float tmin = 0;                       /* :466 */
    orc_node node = VIRTUAL_NODE;         /* :467 */
    hitmin leftRay = hitmin_default(), rightRay = hitmin_default(); /* :468-469 */
    stk_u8_push(&st.A, A_COMPUTE);        /* :470 */
    unsigned char action = A_GOTO_LFT;    /* :471 */
    int run = 1;                          /* :472 */
    uint64_t it = 0;
    while (run || st.A.count > 0) {       /* :474 */
        ++it;
        if (action & A_SAVE_LFT) {        /* :476-481 */
            tmin = stk_f32_pop(&st.T);
            ev->overflow += stk_hit_push(&st.P, leftRay);
            action = A_GOTO_RGH;
        }
        if (action & (A_GOTO_LFT | A_GOTO_RGH)) /* :482-495 */
            go_to(&st, &action, &node, s, &leftRay, &rightRay, ray, &tmin, &run, ev);
        if (action & (A_LOAD_LFT | A_LOAD_RGH | A_COMPUTE)) /* :496-508 */
            compute(&st, &action, &node, s, &leftRay, &rightRay, &tmin, &run, ev);
        if (it > 100000000ull) break;     /* safety net for malformed trees; never reached on valid input */
    }
    ev->iters += it;
    return leftRay;                       /* :511 */
}

/* ------------------------------------------------------------------ hit details */
/* sphereHitDetails :183-200, cylinderHitDetails :338-372, cubeHitDetails :436-457 */
static rayhit hit_details(const orc_scene* s, const ray_t* ray, const hitmin* h)
{
    rayhit d;
    memset(&d, 0, sizeof d);
    d.hit = 0;                            /* RaycastKernel :36 */
    d.t = INFINITY;
    if (h->hit == R_MISS) return d;       /* :37 (hit != Miss) */
    /* note the reference indexes tree.primitives by hitInfo.primitiveIdx (a short) */
    const orc_prim* p = &s->prims[h->primitiveIdx];
    if (h->primitiveType != ORC_SPHERE && h->primitiveType != ORC_CYLINDER && h->primitiveType != ORC_CUBE) return d;
    d.hit = 1;
    d.t = h->t;
    d.position = ray_at(ray, d.t);
    d.primitiveIdx = h->primitiveIdx;
    f3 C = {p->x, p->y, p->z};
    if (h->primitiveType == ORC_SPHERE) {
        d.normal = normalize3(sub3(d.position, C)); /* :189-195; normalize = a * invLen */
    } else if (h->primitiveType == ORC_CUBE) {
        f3 PC = sub3(d.position, C);      /* :446 */
        const float bias = 1.00001f;
        float hs = p->p[0] / 2;
        f3 n = {trunc_via_int((PC.x / hs) * bias), trunc_via_int((PC.y / hs) * bias), trunc_via_int((PC.z / hs) * bias)}; /* :449 */
        d.normal = normalize3(n);         /* :451 */
    } else {
        const f3 V = {p->p[2], p->p[3], p->p[4]};
        const float hh = p->p[1] * 0.5f;
        /* :347 — in this function the reference kernel's SASS forms C with one rounding (FFMA -V, h/2, centre) */
        const f3 Cb = {fmaf(-V.x, hh, p->x), fmaf(-V.y, hh, p->y), fmaf(-V.z, hh, p->z)};
        const f3 OC = sub3(ray->origin, Cb);
        if (h->hit & R_FLAG1) d.normal = neg3(V);       /* :351-355 */
        else if (h->hit & R_FLAG2) d.normal = V;        /* :356-360 */
        else {                                          /* :361-365 */
            float dV = dot3(V, ray->direction);
            float OCV = fmaf(V.z, OC.z, fmaf(V.x, OC.x, V.y * OC.y));
            float m = fmaf(h->t, dV, OCV);
            f3 q = {fmaf(-V.x, m, d.position.x - Cb.x), fmaf(-V.y, m, d.position.y - Cb.y), fmaf(-V.z, m, d.position.z - Cb.z)};
            d.normal = normalize3(q);
        }
    }
    if (h->hit & R_FLIP) d.normal = neg3(d.normal);     /* :196-199 etc. */
    if (h->hit & R_EXIT) d.normal = neg3(d.normal);
    return d;
}

/* ------------------------------------------------------------------ shading */
/* LightningKernel, RaycastingKernels.cu:49-111 */
static void shade(const orc_scene* s, const orc_camera* cam, const rayhit* hi, const float lightDir[3], float out[4])
{
    if (!hi->hit) { out[0] = 0.08f; out[1] = 0.08f; out[2] = 0.11f; out[3] = 1; return; } /* :109 */
    const orc_prim* p = &s->prims[hi->primitiveIdx];
    f3 L = {lightDir[0], lightDir[1], lightDir[2]};
    L = normalize3(L);                                  /* :78 */
    f3 eye = {cam->x, cam->y, cam->z};
    f3 Vv = normalize3(sub3(eye, hi->position));        /* :79 */
    /* reflect(-L, normal), Float3Utils.cuh:39-44: n = normalize(b); a - 2*dot(a,n)*n */
    f3 n = normalize3(hi->normal);
    /* SASS of LightningKernel: FMUL Lx*nx; FFMA -Ly,ny,-that; FFMA -Lz,nz,. (the x product is the one rounded first here) */
    float dn = fmaf(-L.z, n.z, fmaf(-L.y, n.y, -(L.x * n.x)));
    float two = dn + dn;
    f3 R = {fmaf(-n.x, two, -L.x), fmaf(-n.y, two, -L.y), fmaf(-n.z, two, -L.z)};
    float diff = fmaxf(dot3(hi->normal, L), 0.0f);      /* :86 */
    float sb = fmaxf(dot3(Vv, R), 0.0f);                /* :90 */
    float spec = powf(sb, 30.0f);
    float k = fmaf(spec, 0.7f, fmaf(diff, 0.8f, 0.2f)); /* :83-98, white light: ka + kd*diff + ks*spec */
    float c[3] = {p->r * k, p->g * k, p->b * k};
    for (int i = 0; i < 3; ++i) out[i] = fminf(fmaxf(c[i], 0.0f), 1.0f); /* :101-103 */
    out[3] = 1.0f;
}

/* ------------------------------------------------------------------ public entry points */
int orc_render(const orc_scene* s, int w, int h, const orc_camera* cam, const float light_dir[3],
               float tan_half_fov, int y0, int y1, int nthreads,
               uint8_t* hit, int32_t* prim, float* t, uint8_t* flags, float* rgba,
               orc_counters* counters)
{
    if (!s || s->n_nodes <= 0) return 1;
    if (y0 < 0) y0 = 0;
    if (y1 > h) y1 = h;
    /* RaycastKernel :15-16: tan(cam.fov / 2.0f) (CUDA tanf on the device) */
    const float th = tan_half_fov == tan_half_fov ? tan_half_fov : tanf(cam->fov * 0.5f);
    orc_counters tot;
    memset(&tot, 0, sizeof tot);
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        evt ev;
        memset(&ev, 0, sizeof ev);
        uint64_t rays = 0, hits = 0, max_it = 0;
#pragma omp for schedule(dynamic, 8)
        for (int y = y0; y < y1; ++y) {
            for (int x = 0; x < w; ++x) {
                ray_t ray = raygen(cam, (float)w, (float)h, x, y, th);
                uint64_t before = ev.iters;
                hitmin res = csg_ray_cast(s, &ray, &ev);   /* RaycastKernel :32 */
                if (ev.iters - before > max_it) max_it = ev.iters - before;
                rayhit det = hit_details(s, &ray, &res);   /* :35-45 */
                size_t idx = (size_t)y * w + x;            /* :33 */
                rays++;
                hits += det.hit ? 1 : 0;
                if (hit) hit[idx] = det.hit ? 1 : 0;
                if (prim) prim[idx] = det.hit ? det.primitiveIdx : -1;
                if (t) t[idx] = det.hit ? det.t : -1.0f;
                if (flags) flags[idx] = res.hit;
                if (rgba) shade(s, cam, &det, light_dir, rgba + 4 * idx);
            }
        }
#pragma omp critical
        {
            tot.rays += rays; tot.hits += hits; tot.iters += ev.iters; tot.goto_calls += ev.goto_calls;
            tot.compute_calls += ev.compute_calls; tot.aabb += ev.aabb; tot.sphere += ev.sphere;
            tot.cylinder += ev.cylinder; tot.cube += ev.cube;
            if (max_it > tot.max_iters) tot.max_iters = max_it;
            if (ev.max_a > tot.max_action_depth) tot.max_action_depth = ev.max_a;
            if (ev.max_p > tot.max_hit_depth) tot.max_hit_depth = ev.max_p;
            if (ev.max_t > tot.max_time_depth) tot.max_time_depth = ev.max_t;
            tot.stack_overflows += ev.overflow;
        }
    }
    if (counters) *counters = tot;
    return 0;
}

int orc_hit_primitive(const orc_prim* p, int type, const float origin[3], const float dir[3], float tmin,
                      float* t_out, int* flags_out)
{
    ray_t r = {{origin[0], origin[1], origin[2]}, {dir[0], dir[1], dir[2]}};
    evt ev;
    memset(&ev, 0, sizeof ev);
    hitmin h = hitmin_default();
    if (type == ORC_SPHERE) sphere_hit(&r, p, &h, tmin, &ev);
    else if (type == ORC_CYLINDER) cylinder_hit(&r, p, &h, tmin, &ev);
    else if (type == ORC_CUBE) cube_hit(&r, p, &h, tmin, &ev);
    if (t_out) *t_out = h.t;
    if (flags_out) *flags_out = h.hit;
    return !(h.hit & R_MISS);
}

int orc_aabb_hit(const float bmin[3], const float bmax[3], const float origin[3], const float dir[3], float tmin)
{
    ray_t r = {{origin[0], origin[1], origin[2]}, {dir[0], dir[1], dir[2]}};
    evt ev;
    memset(&ev, 0, sizeof ev);
    return bvh_hit(&r, bmin, bmax, NULL, tmin, &ev);
}
