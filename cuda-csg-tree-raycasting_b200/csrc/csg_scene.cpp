// Scene text parser, writer, synthetic generator and GPU tree flattening (host side of libcsg_b200).
//
// Behavioural contract = the reference's CSGTree::Parse (RayCasting/CSGTree/CSGTree.cu:5-152):
// same grammar, keywords, argument counts, range checks and error texts.  This file must be compiled
// WITHOUT floating-point contraction (-ffp-contract=off): cylinder axes and leaf boxes have to come out
// bit-identical to the reference's host code.
#include "csg_scene.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace csgb {

// ------------------------------------------------------------------------------------------ parsing
namespace {

struct Tokens {
    std::vector<std::string> t;
    explicit Tokens(const char* text, size_t len)
    {  // split(): operator>> tokens, CSGTree.cu:181-194
        size_t i = 0;
        auto sp = [](unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; };
        while (i < len) {
            while (i < len && sp((unsigned char)text[i])) ++i;
            size_t b = i;
            while (i < len && !sp((unsigned char)text[i])) ++i;
            if (i > b) t.emplace_back(text + b, i - b);
        }
    }
};

// std::stof / std::stod: longest valid prefix; what() is "stof"/"stod" when nothing converts or on range error.
bool to_float(const std::string& s, float& v)
{
    errno = 0;
    char* end = nullptr;
    v = std::strtof(s.c_str(), &end);
    return end != s.c_str() && errno != ERANGE;
}
bool to_double(const std::string& s, double& v)
{
    errno = 0;
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return end != s.c_str() && errno != ERANGE;
}
// color(): std::stoi(hex2, nullptr, 16) / 255, CSGTree.cu:196-200
bool to_color(const std::string& tok, size_t off, float& v)
{
    char two[3] = {tok[off], tok[off + 1], 0};
    char* end = nullptr;
    long x = std::strtol(two, &end, 16);
    if (end == two) return false;
    v = (float)(int)x / 255;
    return true;
}

// Primitive(...) for cylinders, Primitives.h:56-88: rotate (0,1,0) by Y, X, Z Euler angles in double,
// divide by the SQUARED length, store as float.
void cylinder_axis(double rx, double ry, double rz, float* axis)
{
    rx = rx * 0.017453292519943295769236907684886;
    ry = ry * 0.017453292519943295769236907684886;
    rz = rz * 0.017453292519943295769236907684886;
    double ax = -sin(rz) * cos(ry) + sin(ry) * sin(rx) * cos(rz);
    double ay = cos(rx) * cos(rz);
    double az = sin(ry) * sin(rz) + sin(rx) * cos(ry) * cos(rz);
    double len = (ax * ax + ay * ay + az * az);
    ax /= len;
    ay /= len;
    az /= len;
    axis[0] = (float)ax;
    axis[1] = (float)ay;
    axis[2] = (float)az;
}

// BVHNode(Primitive,type), BVHNode.cuh:17-66 — the reference's leaf boxes (sphere +-2r, cylinder +-max(h/2,r), cube exact)
void ref_leaf_box(const RefPrim& p, int type, float* mn, float* mx)
{
    const float c[3] = {p.x, p.y, p.z};
    for (int i = 0; i < 3; ++i) mn[i] = mx[i] = 0;
    if (type == kSphere) {
        float r = p.p[0];
        for (int i = 0; i < 3; ++i) { mn[i] = c[i] - r - r; mx[i] = c[i] + r + r; }
    } else if (type == kCylinder) {
        float m = std::max(p.p[1] / 2, p.p[0]);
        for (int i = 0; i < 3; ++i) { mn[i] = c[i] - m; mx[i] = c[i] + m; }
    } else if (type == kCube) {
        float h = p.p[0] / 2;
        for (int i = 0; i < 3; ++i) { mn[i] = c[i] - h; mx[i] = c[i] + h; }
    }
}

}  // namespace

std::string parse_scene(const char* text, size_t len, Scene& out)
{
    Tokens tk(text, len);
    const std::vector<std::string>& s = tk.t;
    const int n = (int)s.size();
    out.nodes.clear();
    out.prims.clear();
    std::vector<std::pair<int, int>> open;  // (operator node, children seen)
    int nodes = 0, prims = 0;
    for (int i = 0; i < n; ++i) {
        out.nodes.push_back(RefNode{-1, -1, -1, -1, -1, {0, 0, 0}, {0, 0, 0}});
        if (nodes != 0) {
            if (open.empty()) return "Cannot parse";
            auto& top = open.back();
            if (top.second == 0) {
                out.nodes[top.first].left = nodes;
                out.nodes[nodes].parent = top.first;
                top.second++;
            } else {
                out.nodes[top.first].right = nodes;
                out.nodes[nodes].parent = top.first;
                open.pop_back();
            }
        }
        const std::string& kw = s[i];
        int type = kw == "Union" ? kUnion : kw == "Difference" ? kDifference : kw == "Intersection" ? kIntersection
                   : kw == "Sphere" ? kSphere : kw == "Cylinder" ? kCylinder : kw == "Cube" ? kCube : -1;
        if (type < 0) return "Cannot parse - Unrecognized keyword: " + kw;
        out.nodes[nodes].type = type;
        if (type <= kIntersection) {
            open.push_back({nodes, 0});
        } else {
            const int nargs = type == kCylinder ? 9 : 5;
            if (i + nargs >= n) return "Cannot parse - unexpected end of input";  // reference: out-of-range read (UB)
            RefPrim p{};
            p.id = prims;
            out.nodes[nodes].prim = prims;
            if (!to_float(s[i + 1], p.x) || !to_float(s[i + 2], p.y) || !to_float(s[i + 3], p.z)) return "stof";
            const std::string& col = s[i + 4];
            if (col.size() != 6) return "Cannot parse color " + col;
            if (!to_color(col, 0, p.r) || !to_color(col, 2, p.g) || !to_color(col, 4, p.b)) return "stoi";
            if (!to_float(s[i + 5], p.p[0])) return "stof";
            if (type == kCylinder) {
                double rx, ry, rz;
                if (!to_float(s[i + 6], p.p[1])) return "stof";
                if (!to_double(s[i + 7], rx) || !to_double(s[i + 8], ry) || !to_double(s[i + 9], rz)) return "stod";
                if (rx > 360 || rx < 0) return "Invalid roation rotX should be in range [0, 360] deg";  // sic, CSGTree.cu:103-108
                if (ry > 360 || ry < 0) return "Invalid roation rotY should be in range [0, 360] deg";
                if (rz > 360 || rz < 0) return "Invalid roation rotZ should be in range [0, 360] deg";
                cylinder_axis(rx, ry, rz, &p.p[2]);
            }
            out.prims.push_back(p);
            i += nargs;
            prims++;
        }
        nodes++;
    }
    if (nodes != 2 * prims - 1) return "Cannot parse - number of primitives do not match number of nodes";
    // ConstructBVH, CSGTree.cu:154-179: reference boxes (kept for csg_scene_dump parity)
    for (int id = nodes - 1; id >= 0; --id) {
        RefNode& nd = out.nodes[id];
        if (nd.prim != -1) {
            ref_leaf_box(out.prims[nd.prim], nd.type, nd.bmin, nd.bmax);
        } else {
            const RefNode& l = out.nodes[nd.left];
            const RefNode& r = out.nodes[nd.right];
            for (int k = 0; k < 3; ++k) {
                nd.bmin[k] = std::min(l.bmin[k], r.bmin[k]);
                nd.bmax[k] = std::max(l.bmax[k], r.bmax[k]);
            }
        }
    }
    return "";
}

int Scene::depth() const
{
    if (nodes.empty()) return 0;
    std::vector<int> d(nodes.size(), 0);
    int best = 0;
    for (size_t i = 0; i < nodes.size(); ++i) {  // preorder: parents precede children
        const RefNode& n = nodes[i];
        int here = (n.parent >= 0 ? d[n.parent] : 0) + (n.prim == -1 ? 1 : 0);
        d[i] = here;
        best = std::max(best, here);
    }
    return best;
}

// ------------------------------------------------------------------------------------------ writer / generator
namespace {
const char* kw_of(int type)
{
    static const char* k[] = {"Union", "Difference", "Intersection", "Sphere", "Cylinder", "Cube"};
    return k[type];
}
void put_color(std::string& o, const RefPrim& p)
{
    char b[16];
    auto q = [](float c) { int v = (int)std::lround(c * 255.0f); return std::min(255, std::max(0, v)); };
    std::snprintf(b, sizeof b, "%02X%02X%02X", q(p.r), q(p.g), q(p.b));
    o += b;
}
}  // namespace

std::string write_scene(const Scene& s)
{
    // Cylinder Euler angles are not recoverable from the stored axis in general; the writer emits the
    // rotation that reproduces the axis via rotX (about X) then rotZ: axis = (-sinZ, cosX cosZ, sinX cosZ) with rotY = 0.
    std::string o;
    std::function<void(int, int)> rec = [&](int id, int ind) {
        const RefNode& n = s.nodes[id];
        o.append((size_t)ind, '\t');
        o += kw_of(n.type);
        char b[256];
        if (n.prim == -1) {
            o += "\n";
            rec(n.left, ind + 1);
            rec(n.right, ind + 1);
            return;
        }
        const RefPrim& p = s.prims[n.prim];
        std::snprintf(b, sizeof b, " %.9g %.9g %.9g ", p.x, p.y, p.z);
        o += b;
        put_color(o, p);
        if (n.type == kCylinder) {
            double ax = p.p[2], ay = p.p[3], az = p.p[4];
            double l = std::sqrt(ax * ax + ay * ay + az * az);
            if (l > 0) { ax /= l; ay /= l; az /= l; }
            double rz = std::asin(std::min(1.0, std::max(-1.0, -ax)));   // in [-90,90]
            double rx = std::atan2(az, ay);
            if (std::cos(rz) < 0) rx += M_PI;
            auto deg = [](double r) { double d = std::fmod(r * 180.0 / M_PI, 360.0); if (d < 0) d += 360.0; return d; };
            std::snprintf(b, sizeof b, " %.9g %.9g %.9g 0 %.9g\n", p.p[0], p.p[1], deg(rx), deg(rz));
        } else {
            std::snprintf(b, sizeof b, " %.9g\n", p.p[0]);
        }
        o += b;
    };
    if (!s.nodes.empty()) rec(0, 0);
    return o;
}

namespace {
struct Rng {  // splitmix64
    uint64_t s;
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double range(double a, double b) { return a + (b - a) * uni(); }
};
}  // namespace

// balanced tree; leaves uniform in a 40^3 box centred (0,0,-40); 60% spheres r in [0.3,1.5],
// 20% cubes edge in [0.5,2.5], 20% cylinders r in [0.3,1], h in [1,3], rot in [0,360]^3; operators 70% Union,
// 20% Difference, 10% Intersection (Intersection only at the two lowest operator levels).
std::string generate_scene(int n_primitives, uint64_t seed)
{
    Rng rng{seed};
    std::string o;
    std::function<void(int, int)> rec = [&](int count, int ind) {
        char b[256];
        o.append((size_t)ind, '\t');
        if (count == 1) {
            double x = rng.range(-20, 20), y = rng.range(-20, 20), z = rng.range(-60, -20);
            unsigned col = (unsigned)(rng.next() & 0xFFFFFF);
            double k = rng.uni();
            if (k < 0.6)
                std::snprintf(b, sizeof b, "Sphere %.5f %.5f %.5f %06X %.5f\n", x, y, z, col, rng.range(0.3, 1.5));
            else if (k < 0.8)
                std::snprintf(b, sizeof b, "Cube %.5f %.5f %.5f %06X %.5f\n", x, y, z, col, rng.range(0.5, 2.5));
            else
                std::snprintf(b, sizeof b, "Cylinder %.5f %.5f %.5f %06X %.5f %.5f %.3f %.3f %.3f\n", x, y, z, col,
                              rng.range(0.3, 1.0), rng.range(1.0, 3.0), rng.range(0, 360), rng.range(0, 360), rng.range(0, 360));
            o += b;
            return;
        }
        double k = rng.uni();
        const bool low = count <= 4;  // the two lowest operator levels
        const char* op = k < 0.7 ? "Union" : k < 0.9 ? "Difference" : (low ? "Intersection" : "Union");
        o += op;
        o += "\n";
        int l = (count + 1) / 2;
        rec(l, ind + 1);
        rec(count - l, ind + 1);
    };
    if (n_primitives < 1) n_primitives = 1;
    rec(n_primitives, 0);
    return o;
}

// ------------------------------------------------------------------------------------------ flattening
namespace {

struct Box {
    float mn[3], mx[3];
    void grow(const Box& b) { for (int i = 0; i < 3; ++i) { mn[i] = std::min(mn[i], b.mn[i]); mx[i] = std::max(mx[i], b.mx[i]); } }
    double volume() const { double v = 1; for (int i = 0; i < 3; ++i) v *= std::max(0.0, (double)mx[i] - mn[i]); return v; }
    double area() const { double d[3]; for (int i = 0; i < 3; ++i) d[i] = std::max(0.0, (double)mx[i] - mn[i]); return 2 * (d[0] * d[1] + d[1] * d[2] + d[0] * d[2]); }
};

struct Work {  // working tree node
    int type, prim, left, right;
    Box box;
    bool has_cylinder = false;  // some cylinder below: the box only GATES (Q6), it does not bound the hits
    bool pure = false;          // leaf: sphere or cube; operator: Union of two pure subtrees
    int depth = 0;              // operator levels below and including this node (a primitive: 0)
};

// Culling box of a leaf.  Culling boxes only decide whether a subtree is skipped; the proof that any box
// containing these gives results identical to the reference is in DESIGN.md §"Culling contract".
//   sphere  : centre +- r (slightly inflated) — the reference's +-2r is conservative, so may be tightened freely
//   cube    : the cube itself (identical to the reference's leaf box)
//   cylinder: the reference's own NON-conservative leaf box — it gates the primitive, so it must be kept as is (Q6)
Box leaf_cull_box(const RefPrim& p, int type)
{
    Box b;
    const float c[3] = {p.x, p.y, p.z};
    if (type == kSphere) {
        float r = std::fabs(p.p[0]);
        for (int i = 0; i < 3; ++i) {
            float pad = r * 1e-4f + std::fabs(c[i]) * 4e-7f + 1e-30f;
            b.mn[i] = c[i] - r - pad;
            b.mx[i] = c[i] + r + pad;
        }
    } else {
        ref_leaf_box(p, type, b.mn, b.mx);
        for (int i = 0; i < 3; ++i) if (b.mn[i] > b.mx[i]) std::swap(b.mn[i], b.mx[i]);  // negative sizes
    }
    return b;
}

// Smallest positive float a with fl(fl(a / half) * 1.00001f) >= level (RaycastingKernels.cu:422-424 in the reference's own
// operations: IEEE float division and multiplication, both monotonic in a for half > 0), found by bisection over the bit
// patterns.  0 when half is not a positive normal number (the kernel then always evaluates the reference's expression).
float cube_normal_threshold(float half, float level)
{
    if (!(half >= 1.17549435e-38f) || !std::isfinite(half)) return 0.0f;
    auto reaches = [&](uint32_t bits) {
        float a;
        std::memcpy(&a, &bits, 4);
        volatile float q = a / half;          // volatile: one IEEE rounding per operation, no contraction, no excess precision
        volatile float m = q * 1.00001f;
        return m >= level;
    };
    uint32_t lo = 0u, hi = 0x7f800000u;       // +0 .. +inf; reaches(+inf) is true, reaches(+0) is false
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (reaches(mid)) hi = mid; else lo = mid;
    }
    float a;
    std::memcpy(&a, &hi, 4);
    return std::isfinite(a) ? a : 0.0f;
}

constexpr int kFullOccupancyLevels = 13;   // (13 + 3) frames x 16 B x 768 threads + tree copies = one SM's shared memory (csg_render.cu)

struct Builder {
    const Scene& s;
    std::vector<Work> w;
    int slack = 2;              // build_union: levels a rebuilt Union subtree may be taller than the shortest tree over its operands
    explicit Builder(const Scene& sc) : s(sc) {}

    int add(int type, int prim, int l, int r)
    {
        Work n{type, prim, l, r, Box{}, false};
        n.has_cylinder = prim != -1 ? type == kCylinder : (w[l].has_cylinder || w[r].has_cylinder);
        n.pure = prim != -1 ? (type == kSphere || type == kCube) : (type == kUnion && w[l].pure && w[r].pure);
        n.depth = prim != -1 ? 0 : 1 + std::max(w[l].depth, w[r].depth);
        w.push_back(n);
        return (int)w.size() - 1;
    }

    // Copies the parsed tree shape.
    int copy(int id)
    {
        const RefNode& n = s.nodes[id];
        if (n.prim != -1) return add(n.type, n.prim, -1, -1);
        int l = copy(n.left);
        int r = copy(n.right);
        return add(n.type, -1, l, r);
    }

    // Collects the operands of a maximal Union-only subtree.
    void collect_union(int id, std::vector<int>& items)
    {
        const RefNode& n = s.nodes[id];
        if (n.prim == -1 && n.type == kUnion) {
            collect_union(n.left, items);
            collect_union(n.right, items);
        } else {
            items.push_back(id);
        }
    }

    int build_optimized(int id)
    {
        const RefNode& n = s.nodes[id];
        if (n.prim != -1) return add(n.type, n.prim, -1, -1);
        if (n.type != kUnion) {
            int l = build_optimized(n.left);
            int r = build_optimized(n.right);
            return add(n.type, -1, l, r);
        }
        std::vector<int> items;
        collect_union(id, items);
        std::vector<int> built;
        built.reserve(items.size());
        for (int it : items) built.push_back(build_optimized(it));
        for (int b : built) compute_box(b);
        // height budget of the rebuilt subtree: the traversal keeps one frame per operator level in shared memory, so a tall tree
        // costs resident warps (4096-primitive synthetic scene: 30 levels and 12 warps per SM with count-balanced splits, 12.5 ms;
        // 12 levels, 24 warps, 8.9 ms as parsed).  The shortest tree over operands of heights d_i has ceil(log2(sum 2^d_i)) levels
        // (Kraft); `slack` more are allowed so that splits can follow space.
        int budget = 0;
        while (std::ldexp(1.0, budget) < kraft(built, 0, (int)built.size())) ++budget;
        return build_union(built, 0, (int)built.size(), budget + slack);
    }

    // sum of 2^height over the operands v[lo, hi): they fit under `levels` operator levels iff this is at most 2^levels
    double kraft(const std::vector<int>& v, int lo, int hi) const
    {
        double k = 0;
        for (int i = lo; i < hi; ++i) k += std::ldexp(1.0, std::min(w[v[i]].depth, 1000));
        return k;
    }

    // Spatial split of the operands like a BVH build: along the axis of largest centroid extent, as close to the median as the
    // height budget allows (both sides must fit under budget - 1 levels); when no axis has such a position, the operands are dealt
    // out by height instead (tallest first, to the lighter side), which always fits.
    int build_union(std::vector<int>& v, int lo, int hi, int budget)
    {
        if (hi - lo == 1) return v[lo];
        float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};
        auto cen = [&](int id, int ax) { return 0.5f * (w[id].box.mn[ax] + w[id].box.mx[ax]); };
        for (int i = lo; i < hi; ++i)
            for (int a = 0; a < 3; ++a) { cmn[a] = std::min(cmn[a], cen(v[i], a)); cmx[a] = std::max(cmx[a], cen(v[i], a)); }
        int axes[3] = {0, 1, 2};
        std::sort(axes, axes + 3, [&](int a, int b) { const float ea = cmx[a] - cmn[a], eb = cmx[b] - cmn[b]; return ea > eb || (ea == eb && a < b); });
        const double half = std::ldexp(1.0, std::max(budget - 1, 0));
        const int n = hi - lo, want = (lo + hi) / 2;
        int mid = -1;
        std::vector<double> pre((size_t)n + 1);
        for (int t = 0; t < 3 && mid < 0; ++t) {
            const int ax = axes[t];
            std::sort(v.begin() + lo, v.begin() + hi, [&](int a, int b) { float ca = cen(a, ax), cb = cen(b, ax); return ca < cb || (ca == cb && a < b); });
            pre[0] = 0;
            for (int i = 0; i < n; ++i) pre[(size_t)i + 1] = pre[(size_t)i] + std::ldexp(1.0, std::min(w[v[lo + i]].depth, 1000));
            for (int d = 0; d < n && mid < 0; ++d)                       // positions by distance from the median
                for (int sgn = -1; sgn <= 1 && mid < 0; sgn += 2) {
                    const int m = want + sgn * d;
                    if (m <= lo || m >= hi || (d == 0 && sgn == 1)) continue;
                    if (pre[(size_t)(m - lo)] <= half && pre[(size_t)n] - pre[(size_t)(m - lo)] <= half) mid = m;
                }
        }
        if (mid < 0) {
            // by height: tallest first, each to the side that is lighter so far (weights are powers of two: both sides end up
            // within `half` whenever the whole range fits the budget); then the left side's operands first in v
            std::sort(v.begin() + lo, v.begin() + hi, [&](int a, int b) { return w[a].depth > w[b].depth || (w[a].depth == w[b].depth && a < b); });
            std::vector<int> left, right;
            double wl = 0, wr = 0;
            for (int i = lo; i < hi; ++i) {
                const double k = std::ldexp(1.0, std::min(w[v[i]].depth, 1000));
                if (wl <= wr) { left.push_back(v[i]); wl += k; } else { right.push_back(v[i]); wr += k; }
            }
            if (right.empty()) { right.push_back(left.back()); left.pop_back(); }
            std::copy(left.begin(), left.end(), v.begin() + lo);
            std::copy(right.begin(), right.end(), v.begin() + lo + (long)left.size());
            mid = lo + (int)left.size();
        }
        int l = build_union(v, lo, mid, budget - 1);
        int r = build_union(v, mid, hi, budget - 1);
        int id = add(kUnion, -1, l, r);
        compute_box_shallow(id);
        return id;
    }

    void compute_box_shallow(int id)
    {
        Work& n = w[id];
        if (n.prim != -1) { n.box = leaf_cull_box(s.prims[n.prim], n.type); return; }
        const Box& l = w[n.left].box;
        const Box& r = w[n.right].box;
        if (n.type == kUnion) { n.box = l; n.box.grow(r); }
        else if (n.type == kDifference) n.box = l;                       // result is a subset of the left operand
        else n.box = l.volume() <= r.volume() ? l : r;                  // Intersection: subset of both
    }
    void compute_box(int id)
    {
        if (w[id].prim == -1) { compute_box(w[id].left); compute_box(w[id].right); }
        compute_box_shallow(id);
    }
};

}  // namespace

float cube_normal_threshold_of(float half, float level) { return cube_normal_threshold(half, level); }

void flatten(const Scene& s, int optimize, FlatTree& out)
{
    out.nodes.clear();
    out.prims.clear();
    out.depth = 0;
    out.root_is_leaf = false;
    out.root_pure = false;
    out.parent.clear();
    out.leaf_boxes.clear();
    out.subtree_end.clear();
    if (s.nodes.empty()) return;
    // optimize >= 1: Union subtrees rebuilt spatially under a height budget; of the slacks 2, 1, 0 the largest that keeps the whole
    // tree within kFullOccupancyLevels operator levels (up to there the frame kernel keeps 24 warps per SM), else the shortest tree
    Builder b(s);
    int root = -1;
    if (optimize >= 1) {
        for (int slack = 2; slack >= 0; --slack) {
            Builder t(s);
            t.slack = slack;
            const int r = t.build_optimized(0);
            if (root < 0 || t.w[r].depth < b.w[root].depth) { b.w = t.w; b.slack = slack; root = r; }
            if (b.w[root].depth <= kFullOccupancyLevels) break;
        }
    } else {
        root = b.copy(0);
    }
    b.compute_box(root);

    // primitive records
    out.prims.resize(s.prims.size());
    for (size_t i = 0; i < s.prims.size(); ++i) {
        const RefPrim& p = s.prims[i];
        PrimRec& r = out.prims[i];
        std::memset(&r, 0, sizeof r);
        r.color[0] = p.r; r.color[1] = p.g; r.color[2] = p.b;
        r.centre[0] = p.x; r.centre[1] = p.y; r.centre[2] = p.z;
        r.centre[3] = p.p[0];
    }
    for (const RefNode& n : s.nodes) {
        if (n.prim == -1) continue;
        const RefPrim& p = s.prims[n.prim];
        PrimRec& r = out.prims[n.prim];
        r.color[3] = (float)n.type;
        if (n.type == kCube) {
            r.centre[3] = p.p[0] / 2;  // halfSize, RaycastingKernels.cu:423
            // cube normal (:422-424): n = (float)(int)((pc / halfSize) * 1.00001f) is a monotonic, odd step function of pc.
            // base[0], base[1] = the smallest |pc| that gives +-1 and +-2: below base[0] the component is 0, below base[1] it is
            // +-1, and the kernel falls back to the reference's arithmetic beyond (or always, when both are 0).
            r.base[0] = cube_normal_threshold(r.centre[3], 1.0f);
            r.base[1] = cube_normal_threshold(r.centre[3], 2.0f);
        } else if (n.type == kCylinder) {
            const float hh = p.p[1] * 0.5f;  // height / 2
            for (int k = 0; k < 3; ++k) {
                const float c = k == 0 ? p.x : k == 1 ? p.y : p.z;
                r.axis[k] = p.p[2 + k];
                r.haxis[k] = p.p[2 + k] * hh;                // (h/2)*V, one rounding (FMUL in the reference)
                r.base[k] = c - r.haxis[k];                  // C, RaycastingKernels.cu:208
            }
            r.base[3] = p.p[1];
            r.axis[3] = p.p[0];
        }
    }

    // preorder emission
    out.nodes.reserve(b.w.size());
    struct Item { int id, parent, depth; bool is_right; };
    std::vector<int> index(b.w.size(), -1);
    std::function<void(int, int, int)> emit = [&](int id, int parent, int depth) {
        const Work& n = b.w[id];
        const int me = (int)out.nodes.size();
        index[id] = me;
        out.parent.push_back(parent);
        out.subtree_end.push_back((uint32_t)me + 1u);
        out.nodes.push_back(NodeRec{});
        NodeRec& r = out.nodes[me];
        std::memset(&r, 0, sizeof r);
        if (n.prim != -1) {
            const RefPrim& p = s.prims[n.prim];
            if (n.type == kSphere) {
                r.f[0] = p.x; r.f[1] = p.y; r.f[2] = p.z; r.f[3] = p.p[0];
                r.f[4] = p.x; r.f[5] = p.y; r.f[6] = p.z;
            } else {
                Box rb;
                ref_leaf_box(p, n.type, rb.mn, rb.mx);  // cube: lb/rt; cylinder: gating box — both exactly the reference's
                for (int k = 0; k < 3; ++k) { r.f[k] = rb.mn[k]; r.f[3 + k] = rb.mx[k]; }
            }
            r.meta = (uint32_t)n.type | ((uint32_t)n.prim << 8);
            {
                const Box cb = leaf_cull_box(p, n.type);
                float idbits;
                const int32_t me32 = me;
                std::memcpy(&idbits, &me32, 4);
                const float rec[8] = {cb.mn[0], cb.mn[1], cb.mn[2], idbits, cb.mx[0], cb.mx[1], cb.mx[2], 0.0f};
                out.leaf_boxes.insert(out.leaf_boxes.end(), rec, rec + 8);
            }
            return;
        }
        out.depth = std::max(out.depth, depth + 1);
        for (int k = 0; k < 3; ++k) { r.f[k] = n.box.mn[k]; r.f[3 + k] = n.box.mx[k]; }
        int32_t par = parent;
        std::memcpy(&r.f[6], &par, 4);
        emit(n.left, me, depth + 1);
        const int right = (int)out.nodes.size();
        emit(n.right, me, depth + 1);
        out.subtree_end[me] = (uint32_t)out.nodes.size();
        uint32_t meta = (uint32_t)n.type | ((uint32_t)right << 8);
        if (b.w[n.left].prim != -1) meta |= kMetaLeftLeaf;
        if (b.w[n.right].prim != -1) meta |= kMetaRightLeaf;
        if (!n.has_cylinder) meta |= kMetaBounded;
        if (n.pure) meta |= kMetaPure;
        out.nodes[me].meta = meta;
    };
    emit(root, -1, 0);
    out.root_is_leaf = b.w[root].prim != -1;
    out.root_pure = !out.root_is_leaf && b.w[root].pure;
    {
        Box rb = b.w[root].box;
        if (out.root_is_leaf && b.w[root].type == kCylinder) {   // true bounds: sphere of radius sqrt(r^2 + (h/2)^2) about the centre
            const RefPrim& p = s.prims[b.w[root].prim];
            const float rad = std::sqrt(p.p[0] * p.p[0] + 0.25f * p.p[1] * p.p[1]) * 1.001f;
            const float c[3] = {p.x, p.y, p.z};
            for (int k = 0; k < 3; ++k) { rb.mn[k] = c[k] - rad; rb.mx[k] = c[k] + rad; }
        }
        bool ok = true;
        for (int k = 0; k < 3; ++k) {
            ok = ok && std::isfinite(rb.mn[k]) && std::isfinite(rb.mx[k]) && rb.mn[k] <= rb.mx[k];
            const float pad = 1e-4f * (std::fabs(rb.mn[k]) + std::fabs(rb.mx[k])) + 1e-6f;
            out.root_box[k] = rb.mn[k] - pad;
            out.root_box[3 + k] = rb.mx[k] + pad;
        }
        out.root_box_valid = ok;
    }
}

}  // namespace csgb
