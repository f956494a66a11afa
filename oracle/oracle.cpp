// oracle.cpp — CPU restatement of the reference path tracer's algorithm.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product (libbpt.so, the package, the host
// app) links, imports or executes this file. It may be used only by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as the checker
// or the reported CPU baseline — never as the thing measured as "ours" or shipped.
//
// PARITY STATUS: pinned against the reference's own sources, except the driver's traversal.
//   - The integer RNG (pcg, pcg2d, seed, rand) is pinned by KAT-1, derived
//     independently from shaders/common.glsl and re-derived in tests/test_oracle_kat.py.
//   - The scene input (loader semantics, main.cpp:28-58) is pinned by running the reference's
//     own vendored tinyobjloader (oracle/_ref, built from /root/reference where it lies) and
//     committing its output as tests/golden/cornell_scene.json.
//   - Everything the shader text states (seed, camera, sampling, closest-hit, miss, path update,
//     sample mean, running mean in float and rgba8) is pinned BIT FOR BIT against the reference's
//     own shader text compiled as C++ (oracle/_ref/libref_shade.so: shaders/*.glsl|rgen|rchit|rmiss
//     read where they lie, through the mechanical transform oracle/glsl_to_cpp.py, under
//     oracle/glsl_shim.h); tests/test_ref_shade.py compares live in the build container and
//     against the committed golden images tests/golden/ref_shade_*.npz everywhere else.
//   - PARITY UNPINNED, and unpinnable here: the driver's BVH traversal / triangle test behind
//     traceRayEXT (closed source; the Vulkan RT build cannot run in this image).
//     Its contract (closest opaque hit in [tmin,tmax], no culling, barycentrics of v1,v2) is what
//     both sides implement; the shader-text library is given this file's intersector.
//
// What follows which reference lines (paths relative to the reference checkout):
//   pcg / pcg2d / rand            shaders/common.glsl:13-19, 21-31, 33-37
//   seed, camera ray              shaders/raygen.rgen:47-57
//   path loop / update order      shaders/raygen.rgen:62-84
//   createCoordinateSystem        shaders/raygen.rgen:14-21
//   sampleHemisphere (UNIFORM)    shaders/raygen.rgen:23-30
//   sampleDirection               shaders/raygen.rgen:32-39
//   closest hit (position/normal) shaders/closesthit.rchit:43-64
//   miss                          shaders/miss.rmiss:8-12
//   sample mean + running mean    shaders/raygen.rgen:86-90 (rgba8 image: main.cpp:481-484)
//   traceRayEXT semantics         shaders/raygen.rgen:63-75: closest opaque hit in
//                                 [tmin,tmax], no culling (main.cpp:525), barycentrics of v1,v2
//
// Arithmetic: the f32 instance keeps every operation in float, in shader order, and this file
// is compiled with -ffp-contract=off so no FMA is formed. The f64 instance runs the same
// formulas in double (the random numbers stay the float values rand() produces) and is used to
// measure the rounding-noise floor of the parity metric.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- RNG (common.glsl:13-37)
inline uint32_t pcg(uint32_t& state) {
    uint32_t prev = state * 747796405u + 2891336453u;
    uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
    state = prev;
    return (word >> 22u) ^ word;
}
inline void pcg2d(uint32_t& x, uint32_t& y) {
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    x += y * 1664525u;
    y += x * 1664525u;
    x ^= x >> 16u;
    y ^= y >> 16u;
    x += y * 1664525u;
    y += x * 1664525u;
    x ^= x >> 16u;
    y ^= y >> 16u;
}
// float(0xffffffffu) rounds to 2^32, so the scale is exactly 2^-32 and rand() can return 1.0.
inline float randf(uint32_t& seed) {
    uint32_t v = pcg(seed);
    return static_cast<float>(v) * (1.0f / 4294967296.0f);
}
inline uint32_t make_seed(uint32_t px, uint32_t py, uint32_t k) {  // raygen.rgen:47-48
    uint32_t x = px * k, y = py * k;
    pcg2d(x, y);
    return x + y;
}

// ---------------------------------------------------------------- small vector maths
template <class R> struct V3 { R x, y, z; };
template <class R> inline V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R> inline V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> inline V3<R> operator*(V3<R> a, V3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class R> inline V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <class R> inline V3<R> operator*(R s, V3<R> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class R> inline V3<R> operator/(V3<R> a, R s) { return {a.x / s, a.y / s, a.z / s}; }
template <class R> inline V3<R> operator-(V3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R> inline R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> inline V3<R> cross(V3<R> a, V3<R> b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// normalize(v) = v / sqrt(dot(v,v)); the CUDA shade kernel uses the same formula.
template <class R> inline V3<R> normalize(V3<R> a) { return a / std::sqrt(dot(a, a)); }

// ---------------------------------------------------------------- parameters (mirror of bpt_params)
struct orc_params {
    uint32_t width, height, spp_per_frame, max_depth;
    int32_t frame;
    uint32_t tile_y0, tile_rows;
    float cam_origin[3], cam_target[3], sky[3];
    float tmin, tmax;
    uint32_t accum_mode, sampler;
    uint32_t tile_block, tile_nranks, tile_rank;  // interleaved tiling (include/bpt.h)
    uint32_t rr_start_depth, nee;                 // non-parity estimator switches (include/bpt.h)
};

// the image rows a tile covers, in tile-local order (contiguous rows, or round-robin row blocks)
inline uint32_t tile_local_rows(const orc_params& p) {
    if (p.tile_block) return p.height / p.tile_nranks;
    return p.tile_rows ? p.tile_rows : p.height - p.tile_y0;
}
inline uint32_t tile_global_row(const orc_params& p, uint32_t l) {
    if (p.tile_block) return ((l / p.tile_block) * p.tile_nranks + p.tile_rank) * p.tile_block + l % p.tile_block;
    return p.tile_y0 + l;
}

constexpr uint32_t MISS = 0xffffffffu;

// ---------------------------------------------------------------- scene
struct Tri { float v0[3], v1[3], v2[3]; };

struct BNode {  // binary BVH node of the oracle's own (median split) hierarchy
    float lo[3], hi[3];
    uint32_t left;   // internal: left child index, right = left + 1; leaf: first triangle slot
    uint32_t count;  // 0 = internal
};

struct Scene {
    // world-space triangles, instance-major: world prim id = inst * ntris + prim
    std::vector<Tri> tris;
    std::vector<float> faces;  // per *mesh* primitive: Kd.rgb Ke.rgb
    uint32_t ntris_mesh = 0, ninst = 1;
    // BVH (built when tris.size() > brute_threshold)
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;  // slot -> world prim id
    bool has_bvh = false;
    // next-event estimation (bpt_params.nee): the emissive triangles in primitive order, the cumulative
    // distribution of their areas (float(cumulative / total), both summed in double) and the total area
    std::vector<uint32_t> lights;
    std::vector<float> light_cdf;
    float light_area = 0.0f;
};

// Area-proportional table of the emissive triangles; libbpt builds the same table from the same arrays
// (api.cu build_light_table): areas in double from the float vertices, 0.5 * |(v1-v0) x (v2-v0)|.
void build_lights(Scene& s) {
    s.lights.clear(); s.light_cdf.clear(); s.light_area = 0.0f;
    std::vector<double> cum;
    double total = 0.0;
    for (uint32_t i = 0; i < s.tris.size(); ++i) {
        const float* f = &s.faces[6 * size_t(i % s.ntris_mesh)];
        if (f[3] == 0.0f && f[4] == 0.0f && f[5] == 0.0f) continue;
        const Tri& t = s.tris[i];
        const double e1[3] = {double(t.v1[0]) - t.v0[0], double(t.v1[1]) - t.v0[1], double(t.v1[2]) - t.v0[2]};
        const double e2[3] = {double(t.v2[0]) - t.v0[0], double(t.v2[1]) - t.v0[1], double(t.v2[2]) - t.v0[2]};
        const double c[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double area = 0.5 * std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        if (!(area > 0.0)) continue;
        total += area;
        s.lights.push_back(i);
        cum.push_back(total);
    }
    for (double c : cum) s.light_cdf.push_back(float(c / total));
    if (!s.light_cdf.empty()) s.light_cdf.back() = 1.0f;
    s.light_area = float(total);
}

void build_bvh(Scene& s) {
    const uint32_t n = static_cast<uint32_t>(s.tris.size());
    std::vector<float> cen(3 * size_t(n)), lo(3 * size_t(n)), hi(3 * size_t(n));
    for (uint32_t i = 0; i < n; ++i) {
        const Tri& t = s.tris[i];
        for (int a = 0; a < 3; ++a) {
            float mn = std::min(t.v0[a], std::min(t.v1[a], t.v2[a]));
            float mx = std::max(t.v0[a], std::max(t.v1[a], t.v2[a]));
            lo[3 * size_t(i) + a] = mn;
            hi[3 * size_t(i) + a] = mx;
            cen[3 * size_t(i) + a] = 0.5f * (mn + mx);
        }
    }
    s.order.resize(n);
    std::iota(s.order.begin(), s.order.end(), 0u);
    // every leaf holds >= 2 triangles (or is the root), so n + 1 nodes always suffice
    s.nodes.assign(size_t(n) + 1, BNode{});
    std::atomic<uint32_t> nnodes{1};
    struct Job { uint32_t node, first, count; };
    // splits one job: fills its node, returns true and the two child jobs if it became internal
    auto split = [&](const Job& j, Job& a, Job& b) {
        BNode nd{};
        float clo[3], chi[3];
        for (int ax = 0; ax < 3; ++ax) {
            nd.lo[ax] = clo[ax] = std::numeric_limits<float>::max();
            nd.hi[ax] = chi[ax] = -std::numeric_limits<float>::max();
        }
        for (uint32_t k = j.first; k < j.first + j.count; ++k) {
            size_t p = s.order[k];
            for (int ax = 0; ax < 3; ++ax) {
                nd.lo[ax] = std::min(nd.lo[ax], lo[3 * p + ax]);
                nd.hi[ax] = std::max(nd.hi[ax], hi[3 * p + ax]);
                clo[ax] = std::min(clo[ax], cen[3 * p + ax]);
                chi[ax] = std::max(chi[ax], cen[3 * p + ax]);
            }
        }
        int axis = 0;
        float ext = chi[0] - clo[0];
        for (int ax = 1; ax < 3; ++ax)
            if (chi[ax] - clo[ax] > ext) { ext = chi[ax] - clo[ax]; axis = ax; }
        if (j.count <= 4 || !(ext > 0.0f)) {
            nd.left = j.first;
            nd.count = j.count;
            s.nodes[j.node] = nd;
            return false;
        }
        uint32_t mid = j.count / 2;
        std::nth_element(s.order.begin() + j.first, s.order.begin() + j.first + mid,
                         s.order.begin() + j.first + j.count, [&](uint32_t x, uint32_t y) {
                             float cx = cen[3 * size_t(x) + axis], cy = cen[3 * size_t(y) + axis];
                             return cx < cy || (cx == cy && x < y);  // total order: the build is deterministic
                         });
        nd.left = nnodes.fetch_add(2);
        nd.count = 0;
        s.nodes[j.node] = nd;
        a = {nd.left, j.first, mid};
        b = {nd.left + 1, j.first + mid, j.count - mid};
        return true;
    };
    // breadth-first on one thread until there is enough independent work, then one subtree per task
    std::vector<Job> frontier{{0u, 0u, n}};
    const unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());
    while (!frontier.empty() && frontier.size() < 8 * size_t(nthreads) && frontier.front().count > 4096) {
        std::vector<Job> next;
        for (const Job& j : frontier) {
            Job a, b;
            if (split(j, a, b)) { next.push_back(a); next.push_back(b); }
        }
        frontier.swap(next);
    }
    std::atomic<size_t> cursor{0};
    auto worker = [&]() {
        std::vector<Job> stack;
        for (;;) {
            size_t k = cursor.fetch_add(1);
            if (k >= frontier.size()) break;
            stack.push_back(frontier[k]);
            while (!stack.empty()) {
                Job j = stack.back();
                stack.pop_back();
                Job a, b;
                if (split(j, a, b)) { stack.push_back(a); stack.push_back(b); }
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nthreads && i < frontier.size(); ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    s.nodes.resize(nnodes.load());
    s.has_bvh = true;
}

// Moeller-Trumbore in precision R; accepts t in [tmin, tlimit] (closed), u>=0, v>=0, u+v<=1,
// either side (no culling, main.cpp:525).
template <class R>
inline bool tri_hit(const Tri& tr, V3<R> o, V3<R> d, R tmin, R tlimit, R& t, R& u, R& v) {
    V3<R> v0{R(tr.v0[0]), R(tr.v0[1]), R(tr.v0[2])};
    V3<R> e1 = V3<R>{R(tr.v1[0]), R(tr.v1[1]), R(tr.v1[2])} - v0;
    V3<R> e2 = V3<R>{R(tr.v2[0]), R(tr.v2[1]), R(tr.v2[2])} - v0;
    V3<R> p = cross(d, e2);
    R det = dot(e1, p);
    if (det == R(0)) return false;
    R inv = R(1) / det;
    V3<R> s = o - v0;
    R uu = dot(s, p) * inv;
    if (!(uu >= R(0)) || uu > R(1)) return false;
    V3<R> q = cross(s, e1);
    R vv = dot(d, q) * inv;
    if (!(vv >= R(0)) || uu + vv > R(1)) return false;
    R tt = dot(e2, q) * inv;
    if (!(tt >= tmin) || !(tt <= tlimit)) return false;
    t = tt; u = uu; v = vv;
    return true;
}

struct Hit { float t, u, v; uint32_t prim; };
template <class R> struct HitR { R t, u, v; uint32_t prim; };

// Closest hit; equal distances resolve to the lowest world primitive id in both loops
// (the Cornell asset has exact duplicate triangles).
template <class R>
HitR<R> intersect_brute(const Scene& s, V3<R> o, V3<R> d, R tmin, R tmax) {
    HitR<R> h{tmax, 0, 0, MISS};
    const uint32_t n = static_cast<uint32_t>(s.tris.size());
    for (uint32_t i = 0; i < n; ++i) {
        R t, u, v;
        if (tri_hit<R>(s.tris[i], o, d, tmin, h.t, t, u, v) && (t < h.t || h.prim == MISS)) h = {t, u, v, i};
    }
    return h;
}

template <class R>
HitR<R> intersect_bvh(const Scene& s, V3<R> o, V3<R> d, R tmin, R tmax) {
    HitR<R> h{tmax, 0, 0, MISS};
    // slab test in double regardless of R, with a relative pad, so the BVH can never cull a
    // triangle the brute-force loop would accept. Children are visited near-first (the order
    // changes only the work done: ties in t resolve to the lowest primitive id either way).
    const double ox = double(o.x), oy = double(o.y), oz = double(o.z);
    const double ix = 1.0 / double(d.x), iy = 1.0 / double(d.y), iz = 1.0 / double(d.z);
    auto slab = [&](const BNode& nd, double& tn_out) {
        double t0 = (double(nd.lo[0]) - ox) * ix, t1 = (double(nd.hi[0]) - ox) * ix;
        double tn = std::fmin(t0, t1), tf = std::fmax(t0, t1);
        t0 = (double(nd.lo[1]) - oy) * iy; t1 = (double(nd.hi[1]) - oy) * iy;
        tn = std::fmax(tn, std::fmin(t0, t1)); tf = std::fmin(tf, std::fmax(t0, t1));
        t0 = (double(nd.lo[2]) - oz) * iz; t1 = (double(nd.hi[2]) - oz) * iz;
        tn = std::fmax(tn, std::fmin(t0, t1)); tf = std::fmin(tf, std::fmax(t0, t1));
        // fmin/fmax drop NaNs (0 * inf on a degenerate slab), which only widens the interval
        const double pad = 1e-5 * (1.0 + std::fabs(tn) + std::fabs(tf));
        tn_out = tn - pad;
        if (tn - pad > tf + pad) return false;
        if (tf + pad < double(tmin) || tn - pad > double(h.t)) return false;
        return true;
    };
    struct Entry { uint32_t node; double tn; };
    Entry stack[128];
    int sp = 0;
    double tn_root;
    if (slab(s.nodes[0], tn_root)) stack[sp++] = {0u, tn_root};
    while (sp) {
        const Entry e = stack[--sp];
        if (e.tn > double(h.t)) continue;  // culled by a hit found since it was pushed
        const BNode& nd = s.nodes[e.node];
        if (nd.count) {
            for (uint32_t k = nd.left; k < nd.left + nd.count; ++k) {
                uint32_t p = s.order[k];
                R t, u, v;
                if (tri_hit<R>(s.tris[p], o, d, tmin, h.t, t, u, v) && (t < h.t || h.prim == MISS || p < h.prim)) h = {t, u, v, p};
            }
        } else {
            double ta, tb;
            const bool ha = slab(s.nodes[nd.left], ta), hb = slab(s.nodes[nd.left + 1], tb);
            if (ha && hb) {
                if (ta <= tb) { stack[sp++] = {nd.left + 1, tb}; stack[sp++] = {nd.left, ta}; }
                else { stack[sp++] = {nd.left, ta}; stack[sp++] = {nd.left + 1, tb}; }
            } else if (ha) stack[sp++] = {nd.left, ta};
            else if (hb) stack[sp++] = {nd.left + 1, tb};
        }
    }
    return h;
}

template <class R>
inline HitR<R> intersect(const Scene& s, V3<R> o, V3<R> d, R tmin, R tmax, bool brute) {
    return (brute || !s.has_bvh) ? intersect_brute<R>(s, o, d, tmin, tmax) : intersect_bvh<R>(s, o, d, tmin, tmax);
}

// ---------------------------------------------------------------- shading (raygen.rgen:14-39)
template <class R>
inline void coordinate_system(V3<R> N, V3<R>& T, V3<R>& B) {
    if (std::fabs(N.x) > std::fabs(N.y))
        T = V3<R>{N.z, R(0), -N.x} / std::sqrt(N.x * N.x + N.z * N.z);
    else
        T = V3<R>{R(0), -N.z, N.y} / std::sqrt(N.y * N.y + N.z * N.z);
    B = cross(N, T);
}
template <class R> struct Consts;
template <> struct Consts<float> {
    static constexpr float two_pi = 6.2831855f, pi = 3.1415927f, pdf = 0.15915494f;
};
template <> struct Consts<double> {
    static constexpr double two_pi = 2.0 * 3.14159265358979323846, pi = 3.14159265358979323846,
                            pdf = 1.0 / (2.0 * 3.14159265358979323846);
};
template <class R>
inline V3<R> sample_hemisphere(R r1, R r2) {  // uniform: z = r1 (T1)
    R s = std::sqrt(R(1) - r1 * r1);
    return {std::cos(Consts<R>::two_pi * r2) * s, std::sin(Consts<R>::two_pi * r2) * s, r1};
}
template <class R>
inline V3<R> sample_hemisphere_cosine(R r1, R r2) {  // opt-in, not the reference
    R s = std::sqrt(r1);
    return {std::cos(Consts<R>::two_pi * r2) * s, std::sin(Consts<R>::two_pi * r2) * s, std::sqrt(R(1) - r1)};
}
template <class R>
inline V3<R> sample_direction(R r1, R r2, V3<R> n, uint32_t sampler) {
    V3<R> T, B;
    coordinate_system(n, T, B);
    V3<R> d = sampler == 1 ? sample_hemisphere_cosine(r1, r2) : sample_hemisphere(r1, r2);
    return d.x * T + d.y * B + d.z * n;
}

// closesthit.rchit:50-64 for world prim id `wp`
template <class R>
inline void closest_hit(const Scene& s, uint32_t wp, R u, R v, V3<R>& pos, V3<R>& nrm, V3<R>& brdf, V3<R>& emi) {
    const Tri& t = s.tris[wp];
    V3<R> v0{R(t.v0[0]), R(t.v0[1]), R(t.v0[2])}, v1{R(t.v1[0]), R(t.v1[1]), R(t.v1[2])}, v2{R(t.v2[0]), R(t.v2[1]), R(t.v2[2])};
    R b0 = R(1) - u - v;
    pos = v0 * b0 + v1 * u + v2 * v;
    nrm = -normalize(cross(v1 - v0, v2 - v0));
    const float* f = &s.faces[6 * size_t(wp % s.ntris_mesh)];
    brdf = V3<R>{R(f[0]), R(f[1]), R(f[2])} / Consts<R>::pi;
    emi = V3<R>{R(f[3]), R(f[4]), R(f[5])};
}

// primary ray of (px,py), consuming two rands from `seed` (raygen.rgen:51-57)
template <class R>
inline void camera_ray(const orc_params& p, uint32_t px, uint32_t py, uint32_t& seed, V3<R>& o, V3<R>& d) {
    R r1 = R(randf(seed));
    R r2 = R(randf(seed));
    R sx = R(px) + r1, sy = R(py) + r2;
    R ux = sx / R(p.width), uy = sy / R(p.height);
    R dx = ux * R(2) - R(1), dy = uy * R(2) - R(1);
    o = {R(p.cam_origin[0]), R(p.cam_origin[1]), R(p.cam_origin[2])};
    V3<R> target{dx + R(p.cam_target[0]), dy + R(p.cam_target[1]), R(p.cam_target[2])};
    d = normalize(target - o);
}

// one pixel, all samples of one frame -> sum of sample radiances (before /spp)
template <class R>
V3<R> trace_pixel(const Scene& s, const orc_params& p, uint32_t px, uint32_t py, bool brute, uint64_t& rays) {
    V3<R> color{0, 0, 0};
    for (uint32_t sn = 0; sn < p.spp_per_frame; ++sn) {
        uint32_t k = sn + p.spp_per_frame * uint32_t(p.frame) + 1u;
        uint32_t seed = make_seed(px, py, k);
        V3<R> o, d;
        camera_ray<R>(p, px, py, seed, o, d);
        V3<R> w{1, 1, 1};
        R pdf_prev = R(0);  // NEE: solid-angle pdf the current segment's direction was sampled with
        for (uint32_t depth = 0; depth < p.max_depth; ++depth) {
            HitR<R> h = intersect<R>(s, o, d, R(p.tmin), R(p.tmax), brute);
            ++rays;
            if (h.prim == MISS) {  // miss.rmiss:10-11, then raygen.rgen:76 and the break at :81
                color = color + w * V3<R>{R(p.sky[0]), R(p.sky[1]), R(p.sky[2])};
                break;
            }
            V3<R> pos, n, brdf, emi;
            closest_hit<R>(s, h.prim, h.u, h.v, pos, n, brdf, emi);
            if (!(p.nee && depth > 0)) {
                color = color + w * emi;                  // raygen.rgen:76
            } else if (emi.x != R(0) || emi.y != R(0) || emi.z != R(0)) {
                // NEE: a bounce ray found an emitter the previous vertex also sampled by area: balance heuristic between
                // the bounce pdf it was drawn with (pdf_prev) and the area sampler's solid-angle pdf for this point
                R cy = std::fabs(dot(d, n));
                R pl = cy > R(0) && s.light_area > 0.0f ? h.t * h.t / (cy * R(s.light_area)) : R(0);
                color = color + w * emi * (pdf_prev / (pdf_prev + pl));
            }
            if (p.nee && depth + 1 < p.max_depth && !s.lights.empty()) {
                // next-event estimation (include/bpt.h): one area-proportional point on the emissive triangles
                R rs = R(randf(seed)), ra = R(randf(seed)), rb = R(randf(seed));
                size_t lo = 0, hi = s.light_cdf.size() - 1;  // first entry with cdf > rs (the last one if rs == 1)
                while (lo < hi) { size_t mid = (lo + hi) / 2; if (R(s.light_cdf[mid]) > rs) hi = mid; else lo = mid + 1; }
                const uint32_t lp = s.lights[lo];
                const Tri& lt = s.tris[lp];
                V3<R> a{R(lt.v0[0]), R(lt.v0[1]), R(lt.v0[2])}, b{R(lt.v1[0]), R(lt.v1[1]), R(lt.v1[2])}, c{R(lt.v2[0]), R(lt.v2[1]), R(lt.v2[2])};
                R su = std::sqrt(ra);
                R bu = su * (R(1) - rb), bv = su * rb, b0 = R(1) - bu - bv;
                V3<R> y = a * b0 + b * bu + c * bv;
                V3<R> l = y - pos;
                R r2l = dot(l, l);
                if (lp != h.prim && r2l > R(0)) {
                    R rl = std::sqrt(r2l);
                    V3<R> wd = l / rl;
                    V3<R> ny = -normalize(cross(b - a, c - a));
                    R cx = dot(wd, n), cy = std::fabs(dot(wd, ny));
                    if (cx > R(0) && cy > R(0)) {
                        ++rays;
                        HitR<R> sh = intersect<R>(s, pos, wd, R(p.tmin), rl * R(0.999f), brute);
                        if (sh.prim == MISS) {
                            const float* lf = &s.faces[6 * size_t(lp % s.ntris_mesh)];
                            V3<R> ke{R(lf[3]), R(lf[4]), R(lf[5])};
                            // balance heuristic: f / (p_light + p_bounce), both solid-angle pdfs of this direction
                            R pl = r2l / (cy * R(s.light_area));
                            R pb = p.sampler == 1 ? cx / Consts<R>::pi : Consts<R>::pdf;
                            color = color + w * brdf * ke * (cx / (pl + pb));
                        }
                    }
                }
            }
            o = pos;                                      // :77
            R r1 = R(randf(seed));
            R r2 = R(randf(seed));
            d = sample_direction<R>(r1, r2, n, p.sampler);  // :78
            if (p.sampler == 1) {
                w = w * (brdf * Consts<R>::pi);           // cosine pdf cancels the cosine
                pdf_prev = dot(d, n) / Consts<R>::pi;
            } else {
                w = w * (brdf * dot(d, n) / Consts<R>::pdf);  // :79-80
                pdf_prev = Consts<R>::pdf;
            }
            if (p.rr_start_depth && depth + 1 >= p.rr_start_depth) {   // Russian roulette (include/bpt.h)
                R q = std::min(R(1), std::max(w.x, std::max(w.y, w.z)));
                R r3 = R(randf(seed));
                if (!(r3 < q)) break;
                w = w / q;
            }
        }
    }
    return color;
}

inline float unorm8_roundtrip(float x) {
    float c = std::min(std::max(x, 0.0f), 1.0f);
    if (std::isnan(x)) c = 0.0f;
    return std::nearbyint(c * 255.0f) / 255.0f;
}

template <class R>
void render_rows(const Scene& s, const orc_params& p, uint32_t l0, uint32_t l1, bool brute, float* img, uint64_t& rays) {
    for (uint32_t l = l0; l < l1; ++l)
        for (uint32_t x = 0, y = tile_global_row(p, l); x < p.width; ++x) {
            V3<R> c = trace_pixel<R>(s, p, x, y, brute, rays);
            c = c / R(p.spp_per_frame);                  // raygen.rgen:86
            float* px = img + 4 * (size_t(y) * p.width + x);
            R fr = R(p.frame), fr1 = R(p.frame + 1);
            R nw[4] = {(c.x + R(px[0]) * fr) / fr1, (c.y + R(px[1]) * fr) / fr1, (c.z + R(px[2]) * fr) / fr1,
                       (R(1) + R(px[3]) * fr) / fr1};    // :88-89
            for (int ch = 0; ch < 4; ++ch) px[ch] = p.accum_mode == 1 ? unorm8_roundtrip(float(nw[ch])) : float(nw[ch]);
        }
}

}  // namespace

// ================================================================ C API (ctypes)
extern "C" {

uint32_t orc_pcg(uint32_t* state) { return pcg(*state); }
void orc_pcg2d(uint32_t* x, uint32_t* y) { pcg2d(*x, *y); }
uint32_t orc_seed(uint32_t px, uint32_t py, uint32_t k) { return make_seed(px, py, k); }
float orc_rand(uint32_t* seed) { return randf(*seed); }

// xforms may be NULL (one identity instance). Vertices are transformed to world space in f32,
// x' = m00*x + m01*y + m02*z + m03 evaluated left to right without FMA (SURVEY T12 semantics).
void* orc_scene_create(const float* verts, uint32_t nverts, const uint32_t* idx, uint32_t nidx, const float* faces,
                       uint32_t nfaces, const float* xforms, uint32_t ninst, uint32_t brute_threshold) {
    (void)nverts;
    Scene* s = new Scene;
    uint32_t nt = nidx / 3;
    s->ntris_mesh = nt;
    s->ninst = xforms ? ninst : 1;
    s->faces.assign(faces, faces + 6 * size_t(nfaces));
    s->tris.resize(size_t(nt) * s->ninst);
    for (uint32_t i = 0; i < s->ninst; ++i) {
        const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        const float* m = xforms ? xforms + 12 * size_t(i) : ident;
        for (uint32_t t = 0; t < nt; ++t) {
            Tri& tr = s->tris[size_t(i) * nt + t];
            float* dst[3] = {tr.v0, tr.v1, tr.v2};
            for (int c = 0; c < 3; ++c) {
                const float* v = verts + 3 * size_t(idx[3 * size_t(t) + c]);
                if (xforms)
                    for (int r = 0; r < 3; ++r) dst[c][r] = m[4 * r + 0] * v[0] + m[4 * r + 1] * v[1] + m[4 * r + 2] * v[2] + m[4 * r + 3];
                else
                    for (int r = 0; r < 3; ++r) dst[c][r] = v[r];
            }
        }
    }
    if (s->tris.size() > brute_threshold) build_bvh(*s);
    build_lights(*s);
    return s;
}
void orc_scene_destroy(void* s) { delete static_cast<Scene*>(s); }
uint32_t orc_scene_ntris(void* s) { return static_cast<uint32_t>(static_cast<Scene*>(s)->tris.size()); }

// Renders the tile rows of p into img (full W*H*4 float image, read-modify-write running mean).
// precision: 32 or 64. brute: 1 forces the O(N) loop. Returns rays traced.
uint64_t orc_render(void* scene, const orc_params* p, int precision, int brute, int nthreads, float* img) {
    const Scene& s = *static_cast<Scene*>(scene);
    uint32_t y0 = 0, y1 = tile_local_rows(*p);  // tile-local rows
    if (nthreads <= 0) nthreads = int(std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<uint32_t> next{y0};
    std::atomic<uint64_t> total{0};
    auto worker = [&]() {
        uint64_t rays = 0;
        for (;;) {
            uint32_t y = next.fetch_add(4);
            if (y >= y1) break;
            uint32_t ye = std::min(y + 4, y1);
            if (precision == 64) render_rows<double>(s, *p, y, ye, brute != 0, img, rays);
            else render_rows<float>(s, *p, y, ye, brute != 0, img, rays);
        }
        total += rays;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return total.load();
}

// Primary rays of sample `sample_in_frame` for the tile: rays n*8 {o,tmin,d,tmax}, seeds n.
void orc_generate_rays(const orc_params* p, uint32_t sample_in_frame, float* rays, uint32_t* seeds) {
    size_t i = 0;
    for (uint32_t l = 0; l < tile_local_rows(*p); ++l)
        for (uint32_t x = 0, y = tile_global_row(*p, l); x < p->width; ++x, ++i) {
            uint32_t k = sample_in_frame + p->spp_per_frame * uint32_t(p->frame) + 1u;
            uint32_t seed = make_seed(x, y, k);
            V3<float> o, d;
            camera_ray<float>(*p, x, y, seed, o, d);
            float* r = rays + 8 * i;
            r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = p->tmin;
            r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = p->tmax;
            seeds[i] = seed;
        }
}

// Closest hits for n rays. hits: n * {t,u,v (float), prim (u32)}.
void orc_intersect(void* scene, const float* rays, uint32_t n, int precision, int brute, int nthreads, void* hits_out) {
    const Scene& s = *static_cast<Scene*>(scene);
    Hit* hits = static_cast<Hit*>(hits_out);
    if (nthreads <= 0) nthreads = int(std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            uint32_t b = next.fetch_add(1024);
            if (b >= n) break;
            uint32_t e = std::min(b + 1024, n);
            for (uint32_t i = b; i < e; ++i) {
                const float* r = rays + 8 * size_t(i);
                if (precision == 64) {
                    HitR<double> h = intersect<double>(s, {r[0], r[1], r[2]}, {r[4], r[5], r[6]}, double(r[3]), double(r[7]), brute != 0);
                    hits[i] = {float(h.t), float(h.u), float(h.v), h.prim};
                } else {
                    HitR<float> h = intersect<float>(s, {r[0], r[1], r[2]}, {r[4], r[5], r[6]}, r[3], r[7], brute != 0);
                    hits[i] = {h.t, h.u, h.v, h.prim};
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
}

// The oracle's f32 intersector with the callback contract of oracle/glsl_shim.h (ref_intersect_fn): lets the
// reference's compiled shader text (oracle/_ref/libref_shade.so) trace through the oracle's BVH, so that the two
// renders of a big scene differ only where this restatement of the shader text differs from the text.
// user = the scene handle; bit 0 of the handle's alignment is free, so brute force is selected by orc_intersect_cb_brute.
struct RefHitOut { float t, u, v; uint32_t prim; };
void orc_intersect_cb(void* scene, const float* o, float tmin, const float* d, float tmax, RefHitOut* out) {
    const Scene& s = *static_cast<Scene*>(scene);
    HitR<float> h = intersect<float>(s, {o[0], o[1], o[2]}, {d[0], d[1], d[2]}, tmin, tmax, false);
    *out = {h.t, h.u, h.v, h.prim};
}
void orc_intersect_cb_brute(void* scene, const float* o, float tmin, const float* d, float tmax, RefHitOut* out) {
    const Scene& s = *static_cast<Scene*>(scene);
    HitR<float> h = intersect<float>(s, {o[0], o[1], o[2]}, {d[0], d[1], d[2]}, tmin, tmax, true);
    *out = {h.t, h.u, h.v, h.prim};
}

// One shade step for n paths in f32 (stage-level parity of the shade kernel).
// in : hits n*{t,u,v,prim}, weight n*3, seed n
// out: contrib n*3 (= w * emission or w * sky), new_ray n*8, new_weight n*3, new_seed n, alive n (0/1)
void orc_shade(void* scene, const orc_params* p, const void* hits_in, const float* weight, const uint32_t* seed, uint32_t n,
               float* contrib, float* new_ray, float* new_weight, uint32_t* new_seed, uint8_t* alive) {
    const Scene& s = *static_cast<Scene*>(scene);
    const Hit* hits = static_cast<const Hit*>(hits_in);
    for (uint32_t i = 0; i < n; ++i) {
        V3<float> w{weight[3 * i], weight[3 * i + 1], weight[3 * i + 2]};
        uint32_t sd = seed[i];
        if (hits[i].prim == MISS) {
            V3<float> c = w * V3<float>{p->sky[0], p->sky[1], p->sky[2]};
            contrib[3 * i] = c.x; contrib[3 * i + 1] = c.y; contrib[3 * i + 2] = c.z;
            alive[i] = 0;
            for (int k = 0; k < 8; ++k) new_ray[8 * size_t(i) + k] = 0.0f;
            new_weight[3 * i] = w.x; new_weight[3 * i + 1] = w.y; new_weight[3 * i + 2] = w.z;
            new_seed[i] = sd;
            continue;
        }
        V3<float> pos, nr, brdf, emi;
        closest_hit<float>(s, hits[i].prim, hits[i].u, hits[i].v, pos, nr, brdf, emi);
        V3<float> c = w * emi;
        float r1 = randf(sd), r2 = randf(sd);
        V3<float> d = sample_direction<float>(r1, r2, nr, p->sampler);
        if (p->sampler == 1) w = w * (brdf * Consts<float>::pi);
        else w = w * (brdf * dot(d, nr) / Consts<float>::pdf);
        contrib[3 * i] = c.x; contrib[3 * i + 1] = c.y; contrib[3 * i + 2] = c.z;
        float* r = new_ray + 8 * size_t(i);
        r[0] = pos.x; r[1] = pos.y; r[2] = pos.z; r[3] = p->tmin;
        r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = p->tmax;
        new_weight[3 * i] = w.x; new_weight[3 * i + 1] = w.y; new_weight[3 * i + 2] = w.z;
        new_seed[i] = sd;
        alive[i] = 1;
    }
}

// Synthetic triangle soup. Counter-based: value j of triangle i is
// rand-conversion of pcg(seed + (16*i + j) * 0x9E3779B9). `scale` = ntris^(-1/3) is computed
// by the caller in double and rounded to float so every implementation sees the same bits.
// All arithmetic is single mul + single add per value (no FMA).
void orc_soup(uint32_t ntris, uint32_t seed, float scale, float* verts, uint32_t* idx, float* faces) {
    for (uint32_t i = 0; i < ntris; ++i) {
        float f[15];
        for (uint32_t j = 0; j < 15; ++j) {
            uint32_t st = seed + (16u * i + j) * 0x9E3779B9u;
            f[j] = static_cast<float>(pcg(st)) * (1.0f / 4294967296.0f);
        }
        float c[3] = {f[0] * 2.0f - 1.0f, f[1] * 2.0f - 2.0f, f[2] * 2.0f - 1.0f};
        for (int v = 0; v < 3; ++v)
            for (int a = 0; a < 3; ++a) {
                float off = (f[3 + 3 * v + a] * 2.0f - 1.0f) * scale;
                verts[9 * size_t(i) + 3 * v + a] = c[a] + off;
            }
        for (int a = 0; a < 3; ++a) faces[6 * size_t(i) + a] = f[12 + a] * 0.8f + 0.1f;
        bool em = (i % 128u) == 0u;
        faces[6 * size_t(i) + 3] = em ? 17.0f : 0.0f;
        faces[6 * size_t(i) + 4] = em ? 12.0f : 0.0f;
        faces[6 * size_t(i) + 5] = em ? 4.0f : 0.0f;
        idx[3 * size_t(i)] = 3 * i; idx[3 * size_t(i) + 1] = 3 * i + 1; idx[3 * size_t(i) + 2] = 3 * i + 2;
    }
}

// 30-bit Morton code of a point already normalised to [0,1]^3 (10 bits per axis, x most
// significant within each triple) — the key the LBVH builder sorts on.
uint32_t orc_morton30(float x, float y, float z) {
    auto expand = [](uint32_t v) {
        v = (v * 0x00010001u) & 0xFF0000FFu;
        v = (v * 0x00000101u) & 0x0F00F00Fu;
        v = (v * 0x00000011u) & 0xC30C30C3u;
        v = (v * 0x00000005u) & 0x49249249u;
        return v;
    };
    auto q = [](float v) { return uint32_t(std::min(std::max(v * 1024.0f, 0.0f), 1023.0f)); };
    return expand(q(x)) * 4u + expand(q(y)) * 2u + expand(q(z));
}

unsigned orc_hardware_threads(void) { return std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
