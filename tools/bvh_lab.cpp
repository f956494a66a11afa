// bvh_lab.cpp — CPU laboratory for BVH-quality experiments (development tool, not part of the product or the oracle).
//
// Question it answers before any GPU time is spent: how many BVH8 node visits per ray would a restructured binary tree
// save on the synthetic soup? It rebuilds the pipeline of csrc/build.cu on the host in its
// simplest form — soup generator, 30-bit Morton LBVH (Karras split = highest differing key bit), SAH-optimal 8-wide
// collapse (the same dynamic programme and constants as build.cu) — then optionally restructures the BINARY tree
// (tree rotations; SAH rebuild of small subtrees) and counts, for random secondary-like rays, BVH8 nodes visited,
// nodes visited without any child hit, and triangles tested. Child boxes are unquantised and children are visited
// front to back, so absolute counts are a little below the GPU's; the ratio between variants is what matters.
//
//   g++ -O3 -march=native -std=c++17 -pthread tools/bvh_lab.cpp -o /tmp/bvh_lab
//   /tmp/bvh_lab 1000000 200000        builder / traversal variants
//   /tmp/bvh_lab 1000000 100000 simt   warp-level model of the traversal loop and its triangle-step policies
//   /tmp/bvh_lab 1000000 100000 split  triangle pre-splitting (clipped references before the Morton sort)
#include <algorithm>
#include <array>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct V3 { float x, y, z; };
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(V3 p) { const float v[3] = {p.x, p.y, p.z}; for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], v[a]); hi[a] = std::max(hi[a], v[a]); } }
    float half_area() const { const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; }
};
static inline Box merged(const Box& a, const Box& b) { Box r = a; r.grow(b); return r; }

static inline uint32_t pcg(uint32_t& state) {  // shaders/common.glsl:13-19 (same hash the device soup uses)
    state = state * 747796405u + 2891336453u;
    const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

struct Tri { V3 v[3]; };

// ---------------------------------------------------------------- binary tree
// nodes 0..n-2 internal, n-1..2n-2 leaves (leaf n-1+k = k-th triangle in Morton order)
struct Bin {
    uint32_t n = 0;
    std::vector<uint32_t> left, right, count;  // per node (count: triangles below)
    std::vector<Box> box;
    std::vector<uint32_t> prim;                // leaf k -> triangle
    uint32_t root = 0;
    bool is_leaf(uint32_t x) const { return x >= n - 1; }
};

static uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

struct Ref { Box box; uint32_t tri; };
static Bin build_lbvh_refs(const std::vector<Ref>& refs);
static Bin build_lbvh(const std::vector<Tri>& tris) {
    std::vector<Ref> refs(tris.size());
    for (uint32_t i = 0; i < tris.size(); ++i) { for (int k = 0; k < 3; ++k) refs[i].box.grow(tris[i].v[k]); refs[i].tri = i; }
    return build_lbvh_refs(refs);
}

// ---------------------------------------------------------------- triangle pre-splitting (references)
// a triangle becomes up to 2^levels references: its box is halved along its longest axis and the triangle is clipped
// to each half (Sutherland-Hodgman), recursively; a half is not split further once its longest side is below `min_side`
static void clip_poly(std::vector<V3>& poly, int axis, float plane, bool keep_below) {
    std::vector<V3> out;
    const size_t m = poly.size();
    auto get = [&](const V3& p) { return axis == 0 ? p.x : axis == 1 ? p.y : p.z; };
    for (size_t i = 0; i < m; ++i) {
        const V3 a = poly[i], b = poly[(i + 1) % m];
        const float da = get(a) - plane, db = get(b) - plane;
        const bool ia = keep_below ? da <= 0.f : da >= 0.f, ib = keep_below ? db <= 0.f : db >= 0.f;
        if (ia) out.push_back(a);
        if (ia != ib) { const float t = da / (da - db); out.push_back(a + (b - a) * t); }
    }
    poly.swap(out);
}
static void split_rec(const std::vector<V3>& poly, uint32_t tri, int levels, float min_side, std::vector<Ref>& out) {
    Box bx; for (const V3& p : poly) bx.grow(p);
    int ax = 0; float ext = 0.f;
    for (int a = 0; a < 3; ++a) if (bx.hi[a] - bx.lo[a] > ext) { ext = bx.hi[a] - bx.lo[a]; ax = a; }
    if (levels == 0 || ext <= min_side) { out.push_back(Ref{bx, tri}); return; }
    const float mid = 0.5f * (bx.lo[ax] + bx.hi[ax]);
    std::vector<V3> lo = poly, hi = poly;
    clip_poly(lo, ax, mid, true); clip_poly(hi, ax, mid, false);
    if (lo.size() < 3 || hi.size() < 3) { out.push_back(Ref{bx, tri}); return; }
    split_rec(lo, tri, levels - 1, min_side, out); split_rec(hi, tri, levels - 1, min_side, out);
}
static std::vector<Ref> presplit(const std::vector<Tri>& tris, int levels, float min_side) {
    std::vector<Ref> out; out.reserve(tris.size() << levels);
    for (uint32_t i = 0; i < tris.size(); ++i) split_rec({tris[i].v[0], tris[i].v[1], tris[i].v[2]}, i, levels, min_side, out);
    return out;
}

static Bin build_lbvh_refs(const std::vector<Ref>& refs) {
    const uint32_t n = (uint32_t)refs.size();
    std::vector<Box> pb(n);
    Box scene;
    for (uint32_t i = 0; i < n; ++i) { pb[i] = refs[i].box; scene.grow(pb[i]); }
    std::vector<uint64_t> keys(n);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t q[3];
        for (int a = 0; a < 3; ++a) {
            const float c = 0.5f * (pb[i].lo[a] + pb[i].hi[a]);
            const float u = (c - scene.lo[a]) / std::max(scene.hi[a] - scene.lo[a], 1e-30f);
            q[a] = (uint32_t)std::min(std::max(u * 1024.0f, 0.0f), 1023.0f);
        }
        const uint32_t m = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
        keys[i] = ((uint64_t)m << 32) | i;
    }
    std::sort(keys.begin(), keys.end());
    Bin b;
    b.n = n;
    b.left.assign(2 * n - 1, 0); b.right.assign(2 * n - 1, 0); b.count.assign(2 * n - 1, 1); b.box.resize(2 * n - 1);
    b.prim.resize(n);
    for (uint32_t k = 0; k < n; ++k) { const uint32_t ri = (uint32_t)(keys[k] & 0xffffffffu); b.prim[k] = refs[ri].tri; b.box[n - 1 + k] = pb[ri]; }
    // top-down: split a key range at the highest differing bit (what Karras' bottom-up construction yields)
    uint32_t next = 0;
    struct Job { uint32_t lo, hi, node; };
    std::vector<Job> st;
    if (n > 1) st.push_back({0, n - 1, next++});
    while (!st.empty()) {
        const Job j = st.back(); st.pop_back();
        const uint64_t x = keys[j.lo] ^ keys[j.hi];
        const int bit = 63 - __builtin_clzll(x);
        // first index in (lo, hi] whose key has `bit` set
        uint32_t a = j.lo, c = j.hi;
        while (a + 1 < c) { const uint32_t m = (a + c) / 2; if ((keys[m] >> bit) & 1) c = m; else a = m; }
        const uint32_t split = c;  // left = [lo, split-1], right = [split, hi]
        auto child = [&](uint32_t lo, uint32_t hi) -> uint32_t {
            if (lo == hi) return n - 1 + lo;
            const uint32_t id = next++;
            st.push_back({lo, hi, id});
            return id;
        };
        b.left[j.node] = child(j.lo, split - 1);
        b.right[j.node] = child(split, j.hi);
    }
    b.root = 0;
    return b;
}

// boxes and counts of every internal node, post-order (iterative)
static void refit(Bin& b) {
    if (b.n < 2) return;
    std::vector<uint32_t> order; order.reserve(b.n);
    std::vector<uint32_t> st{b.root};
    while (!st.empty()) { const uint32_t x = st.back(); st.pop_back(); if (b.is_leaf(x)) continue; order.push_back(x); st.push_back(b.left[x]); st.push_back(b.right[x]); }
    for (size_t i = order.size(); i-- > 0;) {
        const uint32_t x = order[i];
        b.box[x] = merged(b.box[b.left[x]], b.box[b.right[x]]);
        b.count[x] = b.count[b.left[x]] + b.count[b.right[x]];
    }
}

static double sah_binary(const Bin& b) {
    double c = 0; const double ra = b.box[b.root].half_area();
    for (uint32_t x = 0; x + 1 < b.n; ++x) c += b.box[x].half_area();
    for (uint32_t k = 0; k < b.n; ++k) c += 0.6 * b.box[b.n - 1 + k].half_area();
    return c / ra;
}

// ---------------------------------------------------------------- restructuring 1: tree rotations (Kensler 2008)
static int rotate_pass(Bin& b) {
    std::vector<uint32_t> order; std::vector<uint32_t> st{b.root};
    while (!st.empty()) { const uint32_t x = st.back(); st.pop_back(); if (b.is_leaf(x)) continue; order.push_back(x); st.push_back(b.left[x]); st.push_back(b.right[x]); }
    int done = 0;
    for (size_t i = order.size(); i-- > 0;) {
        const uint32_t x = order[i];
        uint32_t* ch[2] = {&b.left[x], &b.right[x]};
        float best = 0.f; int bs = -1, bg = -1;
        for (int s = 0; s < 2; ++s) {            // swap child ch[s] with a grandchild under ch[1-s]
            const uint32_t o = *ch[1 - s];
            if (b.is_leaf(o)) continue;
            const uint32_t g[2] = {b.left[o], b.right[o]};
            for (int k = 0; k < 2; ++k) {        // grandchild g[k] goes up, ch[s] goes down next to g[1-k]
                const float gain = b.box[o].half_area() - merged(b.box[*ch[s]], b.box[g[1 - k]]).half_area();
                if (gain > best) { best = gain; bs = s; bg = k; }
            }
        }
        if (bs >= 0) {
            const uint32_t o = *ch[1 - bs];
            uint32_t* g = bg == 0 ? &b.left[o] : &b.right[o];
            std::swap(*ch[bs], *g);
            b.box[o] = merged(b.box[b.left[o]], b.box[b.right[o]]);
            b.count[o] = b.count[b.left[o]] + b.count[b.right[o]];
            ++done;
        }
        b.box[x] = merged(b.box[b.left[x]], b.box[b.right[x]]);
        b.count[x] = b.count[b.left[x]] + b.count[b.right[x]];
    }
    return done;
}

// ---------------------------------------------------------------- restructuring 2: SAH rebuild of small subtrees
// every maximal subtree with at most `max_leaves` triangles gets a new topology: top-down full-sweep SAH on the
// primitive-box centroids; the internal node ids of the subtree are reused
static void rebuild_small_subtrees(Bin& b, uint32_t max_leaves) {
    std::vector<uint32_t> st{b.root};
    std::vector<uint32_t> leaves, ids;
    while (!st.empty()) {
        const uint32_t x = st.back(); st.pop_back();
        if (b.is_leaf(x)) continue;
        if (b.count[x] > max_leaves) { st.push_back(b.left[x]); st.push_back(b.right[x]); continue; }
        leaves.clear(); ids.clear();
        std::vector<uint32_t> w{x};
        while (!w.empty()) { const uint32_t y = w.back(); w.pop_back(); if (b.is_leaf(y)) leaves.push_back(y); else { ids.push_back(y); w.push_back(b.left[y]); w.push_back(b.right[y]); } }
        size_t next_id = 0;
        std::function<uint32_t(uint32_t*, uint32_t, bool)> build = [&](uint32_t* L, uint32_t m, bool is_root) -> uint32_t {
            if (m == 1) return L[0];
            const uint32_t id = is_root ? x : ids[next_id++];
            float best = FLT_MAX; int ba = 0; uint32_t bk = m / 2;
            std::vector<float> ra(m);
            for (int a = 0; a < 3; ++a) {
                std::sort(L, L + m, [&](uint32_t p, uint32_t q) { return b.box[p].lo[a] + b.box[p].hi[a] < b.box[q].lo[a] + b.box[q].hi[a]; });
                Box acc;
                for (uint32_t i = m; i-- > 1;) { acc.grow(b.box[L[i]]); ra[i] = acc.half_area(); }
                Box la;
                for (uint32_t i = 1; i < m; ++i) {
                    la.grow(b.box[L[i - 1]]);
                    const float c = la.half_area() * i + ra[i] * (m - i);
                    if (c < best) { best = c; ba = a; bk = i; }
                }
            }
            std::sort(L, L + m, [&](uint32_t p, uint32_t q) { return b.box[p].lo[ba] + b.box[p].hi[ba] < b.box[q].lo[ba] + b.box[q].hi[ba]; });
            const uint32_t l = build(L, bk, false), r = build(L + bk, m - bk, false);
            b.left[id] = l; b.right[id] = r;
            b.box[id] = merged(b.box[l], b.box[r]);
            b.count[id] = b.count[l] + b.count[r];
            return id;
        };
        // ids[0] == x (first popped); hand out the others
        next_id = 1;
        build(leaves.data(), (uint32_t)leaves.size(), true);
    }
    refit(b);
}

// ---------------------------------------------------------------- SAH-optimal collapse (build.cu dp_leaf / dp_internal)
constexpr uint32_t kMaxLeafTris = 2;
constexpr float kCostNode = 1.0f, kCostPrim = 0.6f;
struct Wide {
    struct Node { Box cb[8]; int32_t child[8]; uint8_t ntri[8]; uint32_t tri[8][2]; int n = 0; uint8_t slot[8]; };  // child >= 0: internal node
    std::vector<Node> nodes;
};

static Wide collapse(const Bin& b) {
    const uint32_t N = 2 * b.n - 1;
    std::vector<std::array<float, 8>> c(N);      // c[x][i], i = 1..7
    std::vector<std::array<uint8_t, 8>> d(N);
    std::vector<uint32_t> order; std::vector<uint32_t> st{b.root};
    while (!st.empty()) { const uint32_t x = st.back(); st.pop_back(); order.push_back(x); if (!b.is_leaf(x)) { st.push_back(b.left[x]); st.push_back(b.right[x]); } }
    for (size_t oi = order.size(); oi-- > 0;) {
        const uint32_t x = order[oi];
        const float area = b.box[x].half_area();
        if (b.is_leaf(x)) { for (int i = 1; i <= 7; ++i) c[x][i] = area * kCostPrim; d[x][0] = 1; for (int j = 1; j < 8; ++j) d[x][j] = 0x80; continue; }
        const auto& cl = c[b.left[x]]; const auto& cr = c[b.right[x]];
        float cd[9]; uint8_t kb[9];
        for (int j = 2; j <= 8; ++j) {
            float best = FLT_MAX; int bk = 1;
            for (int k = 1; k < j; ++k) { if (k > 7 || j - k > 7) continue; const float v = cl[k] + cr[j - k]; if (v < best) { best = v; bk = k; } }
            cd[j] = best; kb[j] = (uint8_t)bk;
        }
        const bool leaf = b.count[x] <= kMaxLeafTris;
        c[x][1] = leaf ? area * (float)b.count[x] * kCostPrim : cd[8] + area * kCostNode;
        d[x][0] = leaf ? 1 : 0; d[x][7] = kb[8];
        for (int i = 2; i <= 7; ++i) {
            if (cd[i] < c[x][i - 1]) { c[x][i] = cd[i]; d[x][i - 1] = kb[i]; }
            else { c[x][i] = c[x][i - 1]; d[x][i - 1] = (uint8_t)(kb[i] | 0x80); }
        }
    }
    Wide w;
    // top-down: node job = binary node whose subtree fills one BVH8 node
    struct Job { uint32_t bin; int32_t wide; };
    std::vector<Job> jobs;
    w.nodes.emplace_back();
    jobs.push_back({b.root, 0});
    while (!jobs.empty()) {
        const Job j = jobs.back(); jobs.pop_back();
        // distribute bin over 8 slots
        std::vector<std::pair<uint32_t, int>> work{{j.bin, 8}};
        std::vector<uint32_t> kids;
        bool first = true;
        while (!work.empty()) {
            auto [x, slots] = work.back(); work.pop_back();
            if (!first && slots == 1) { kids.push_back(x); continue; }
            if (b.is_leaf(x)) { kids.push_back(x); continue; }
            int s = slots;
            if (!first) while (s >= 2 && (d[x][s - 1] & 0x80)) --s;
            if (!first && s == 1) { kids.push_back(x); continue; }
            const int k = d[x][s - 1] & 0x7f;
            first = false;
            work.push_back({b.right[x], s - k});
            work.push_back({b.left[x], k});
        }
        for (uint32_t x : kids) {
            Wide::Node& nd = w.nodes[j.wide];
            const int s = nd.n;
            if (s >= 8) { fprintf(stderr, "collapse overflow\n"); exit(1); }
            w.nodes[j.wide].cb[s] = b.box[x];
            if (b.is_leaf(x) || d[x][0] == 1) {
                std::vector<uint32_t> ls, q{x};
                while (!q.empty()) { const uint32_t y = q.back(); q.pop_back(); if (b.is_leaf(y)) ls.push_back(b.prim[y - (b.n - 1)]); else { q.push_back(b.left[y]); q.push_back(b.right[y]); } }
                w.nodes[j.wide].child[s] = -1;
                w.nodes[j.wide].ntri[s] = (uint8_t)ls.size();
                for (size_t t = 0; t < ls.size() && t < 2; ++t) w.nodes[j.wide].tri[s][t] = ls[t];
                w.nodes[j.wide].n = s + 1;
            } else {
                const int32_t id = (int32_t)w.nodes.size();
                w.nodes[j.wide].child[s] = id;
                w.nodes[j.wide].ntri[s] = 0;
                w.nodes[j.wide].n = s + 1;
                w.nodes.emplace_back();
                jobs.push_back({x, id});
            }
        }
    }
    return w;
}

// ---------------------------------------------------------------- octant slots (build.cu k_bvh8_collapse)
// slot bit 2/1/0 set = child on the +x/+y/+z side of the node centre; greedy maximum of dot(sign(slot), child centre -
// node centre) over unassigned pairs. The traversal visits slot s with priority s ^ oct, highest first.
static void assign_slots(Wide& w) {
    for (auto& nd : w.nodes) {
        Box nb;
        for (int c = 0; c < nd.n; ++c) nb.grow(nd.cb[c]);
        float d[8][3];
        for (int c = 0; c < nd.n; ++c)
            for (int a = 0; a < 3; ++a) d[c][a] = 0.5f * (nd.cb[c].lo[a] + nd.cb[c].hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
        bool cu[8] = {}, su[8] = {};
        for (int it = 0; it < nd.n; ++it) {
            float best = -FLT_MAX; int bc = 0, bs = 0;
            for (int c = 0; c < nd.n; ++c) if (!cu[c])
                for (int sl = 0; sl < 8; ++sl) if (!su[sl]) {
                    const float v = ((sl & 4) ? d[c][0] : -d[c][0]) + ((sl & 2) ? d[c][1] : -d[c][1]) + ((sl & 1) ? d[c][2] : -d[c][2]);
                    if (v > best) { best = v; bc = c; bs = sl; }
                }
            cu[bc] = su[bs] = true; nd.slot[bc] = (uint8_t)bs;
        }
    }
}

// ---------------------------------------------------------------- 8-bit child boxes (build.cu k_bvh8_collapse)
// mode 1: one power-of-two step per node (what the 64-byte node stores); mode 2: one per axis (Ylitie's 80-byte node)
static void quantise(Wide& w, int mode) {
    for (auto& nd : w.nodes) {
        Box nb;
        for (int c = 0; c < nd.n; ++c) nb.grow(nd.cb[c]);
        float step[3];
        float mx = 0.f;
        for (int a = 0; a < 3; ++a) mx = std::max(mx, nb.hi[a] - nb.lo[a]);
        for (int a = 0; a < 3; ++a) {
            const float ext = mode == 1 ? mx : nb.hi[a] - nb.lo[a];
            int k = -100;
            if (ext > 0.f) std::frexp(ext * (1.02f / 255.0f), &k);
            step[a] = std::ldexp(1.0f, k);
        }
        for (int c = 0; c < nd.n; ++c)
            for (int a = 0; a < 3; ++a) {
                const float m = 0.0078125f * step[a];
                nd.cb[c].lo[a] = nb.lo[a] + std::floor((nd.cb[c].lo[a] - m - nb.lo[a]) / step[a]) * step[a];
                nd.cb[c].hi[a] = nb.lo[a] + std::ceil((nd.cb[c].hi[a] + m - nb.lo[a]) / step[a]) * step[a];
            }
    }
}

// ---------------------------------------------------------------- traversal statistics
struct Stats { double nodes = 0, empty = 0, tris = 0, hits = 0; };

// order 0: children front to back by entry distance, stale stack entries skipped (ideal); 1: same order, stale
// entries visited (they find no child: what the GPU does); 2: octant order, stale entries visited (the GPU kernel)
static Stats trace_stats(const Wide& w, const std::vector<Tri>& tris, uint32_t nrays, uint32_t seed, int order) {
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    std::vector<Stats> part(T);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] {
        Stats s;
        for (uint32_t r = t; r < nrays; r += T) {
            uint32_t st = seed + r * 0x9E3779B9u;
            auto rnd = [&] { return (float)pcg(st) * 2.3283064365386963e-10f; };
            // origin on a random triangle, direction uniform on the sphere: what a diffuse bounce in the soup looks like
            const Tri& tr = tris[pcg(st) % tris.size()];
            float u = rnd(), v = rnd(); if (u + v > 1.f) { u = 1.f - u; v = 1.f - v; }
            const V3 o = tr.v[0] + (tr.v[1] - tr.v[0]) * u + (tr.v[2] - tr.v[0]) * v;
            const float z = 2.f * rnd() - 1.f, ph = 6.2831853f * rnd(), rr = std::sqrt(std::max(0.f, 1.f - z * z));
            const V3 dir{rr * std::cos(ph), rr * std::sin(ph), z};
            const float inv[3] = {1.f / dir.x, 1.f / dir.y, 1.f / dir.z}, org[3] = {o.x, o.y, o.z};
            const float tmin = 1e-4f;
            float tbest = 1e30f;
            int32_t stack[128]; float stn[128]; int sp = 0;
            stack[sp] = 0; stn[sp++] = 0.f;
            while (sp) {
                const int32_t ni = stack[--sp];
                if ((order == 0 || order >= 3) && stn[sp] > tbest) continue;
                const Wide::Node& nd = w.nodes[ni];
                s.nodes += 1;
                int hit_n = 0; int hs[8]; float ht[8];
                for (int c = 0; c < nd.n; ++c) {
                    float tn = tmin, tf = tbest;
                    for (int a = 0; a < 3; ++a) {
                        float t0 = (nd.cb[c].lo[a] - org[a]) * inv[a], t1 = (nd.cb[c].hi[a] - org[a]) * inv[a];
                        if (t0 > t1) std::swap(t0, t1);
                        tn = std::max(tn, t0); tf = std::min(tf, t1);
                    }
                    if (tn <= tf) { hs[hit_n] = c; ht[hit_n++] = tn; }
                }
                if (!hit_n) { s.empty += 1; continue; }
                // triangles first (they shorten the ray), then internal children far to near onto the stack
                for (int h = 0; h < hit_n; ++h) {
                    const int c = hs[h];
                    if (nd.child[c] >= 0) continue;
                    for (int k = 0; k < nd.ntri[c]; ++k) {
                        s.tris += 1;
                        const Tri& q = tris[nd.tri[c][k]];
                        const V3 e1 = q.v[1] - q.v[0], e2 = q.v[2] - q.v[0], p = cross(dir, e2);
                        const float det = dot(e1, p);
                        if (det == 0.f) continue;
                        const float id = 1.f / det; const V3 sv = o - q.v[0];
                        const float uu = dot(sv, p) * id; const V3 qq = cross(sv, e1);
                        const float vv = dot(dir, qq) * id, tt = dot(e2, qq) * id;
                        if (uu >= 0.f && vv >= 0.f && uu + vv <= 1.f && tt >= tmin && tt < tbest) tbest = tt;
                    }
                }
                int idx[8], m = 0;
                for (int h = 0; h < hit_n; ++h) if (nd.child[hs[h]] >= 0) idx[m++] = h;
                if (order < 2) std::sort(idx, idx + m, [&](int a, int c) { return ht[a] > ht[c]; });
                else {
                    const int oct = (dir.x >= 0.f ? 4 : 0) | (dir.y >= 0.f ? 2 : 0) | (dir.z >= 0.f ? 1 : 0);
                    std::sort(idx, idx + m, [&](int a, int c) { return (nd.slot[hs[a]] ^ oct) < (nd.slot[hs[c]] ^ oct); });  // lowest priority first onto the stack
                }
                // order 3: the bound a GPU stack entry could carry cheaply = min entry distance over the siblings that
                // are pushed as one group (everything but the child descended first); order 4: exact per child
                float gmin = FLT_MAX;
                for (int i = 0; i + 1 < m; ++i) gmin = std::min(gmin, ht[idx[i]]);
                for (int i = 0; i < m; ++i) {
                    stack[sp] = nd.child[hs[idx[i]]];
                    stn[sp++] = order == 3 ? (i + 1 < m ? gmin : -1.f) : ht[idx[i]];
                }
            }
            if (tbest < 1e29f) s.hits += 1;
        }
        part[t] = s;
    });
    for (auto& x : th) x.join();
    Stats s;
    for (auto& p : part) { s.nodes += p.nodes; s.empty += p.empty; s.tris += p.tris; s.hits += p.hits; }
    s.nodes /= nrays; s.empty /= nrays; s.tris /= nrays; s.hits /= nrays;
    return s;
}

// ---------------------------------------------------------------- warp-level model of k_trace's loop (trace.cu)
// 32 lanes in lockstep: per iteration (pop) -> node step -> triangle step -> (terminate); a lane does at most one node
// step and `tris_per_step` triangle tests per iteration, parks a draining triangle group on its stack when a node
// produces a new one, and idle lanes are refilled when fewer than `refill_below` are live (voted every
// `steps_per_refill` iterations). A phase costs its issue slots once per warp, however many lanes take part; the
// per-phase instruction counts are the ncu source-view shares of profiles/r1f (375 per iteration: node step incl.
// pick/push 240, triangle step 74, pop + refill + terminate + loop 63). Policies decide WHEN the triangle step runs.
struct Policy {
    const char* name;
    int tri_every = 1;       // triangle step only every n-th iteration
    int tri_min_lanes = 1;   // ... and only if this many lanes have a triangle pending, or a lane has nothing else to do
    int tris_per_step = 1;
    int node_steps = 1;      // (pop + node step) repetitions per iteration
    bool cull_stale = false; // pushed sibling groups carry the minimum entry distance of their members
};
struct Ray { float o[3], d[3], inv[3], tmin, tbest; int oct; };
struct NGroup { int32_t node = -1; uint8_t mask = 0; float bound = -1.f; };   // remaining hit internal children of `node` (bit = child index)
struct TGroup { uint8_t n = 0; uint32_t tri[16]; };
struct SEntry { bool is_tri; NGroup g; TGroup t; };
struct Lane { bool active = false; bool fresh = false; Ray r; NGroup G; TGroup T; std::vector<SEntry> st; };

static Ray make_ray(const std::vector<Tri>& tris, uint32_t r, uint32_t seed) {
    uint32_t st = seed + r * 0x9E3779B9u;
    auto rnd = [&] { return (float)pcg(st) * 2.3283064365386963e-10f; };
    const Tri& tr = tris[pcg(st) % tris.size()];
    float u = rnd(), v = rnd(); if (u + v > 1.f) { u = 1.f - u; v = 1.f - v; }
    const V3 o = tr.v[0] + (tr.v[1] - tr.v[0]) * u + (tr.v[2] - tr.v[0]) * v;
    const float z = 2.f * rnd() - 1.f, ph = 6.2831853f * rnd(), rr = std::sqrt(std::max(0.f, 1.f - z * z));
    Ray q;
    q.o[0] = o.x; q.o[1] = o.y; q.o[2] = o.z;
    q.d[0] = rr * std::cos(ph); q.d[1] = rr * std::sin(ph); q.d[2] = z;
    for (int a = 0; a < 3; ++a) q.inv[a] = 1.f / q.d[a];
    q.tmin = 1e-4f; q.tbest = 1e30f;
    q.oct = (q.d[0] >= 0.f ? 4 : 0) | (q.d[1] >= 0.f ? 2 : 0) | (q.d[2] >= 0.f ? 1 : 0);
    return q;
}

struct SimOut { double warp_iters = 0, lane_iters = 0, wnode = 0, wtri = 0, nodes = 0, tris = 0, cost = 0; };

static SimOut simulate(const Wide& w, const std::vector<Tri>& tris, uint32_t nrays, const Policy& pol, int refill_below = 30,
                       int steps_per_refill = 2) {
    const unsigned T = std::max(1u, std::thread::hardware_concurrency());
    const uint32_t nwarps = 64;  // independent warps, each pulling 32-ray pools from its own range
    std::vector<SimOut> part(nwarps);
    std::atomic<uint32_t> next_warp{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&] {
        for (;;) {
            const uint32_t wi = next_warp.fetch_add(1);
            if (wi >= nwarps) break;
            SimOut s;
            uint32_t ray_lo = (uint64_t)nrays * wi / nwarps, ray_hi = (uint64_t)nrays * (wi + 1) / nwarps;
            Lane lane[32];
            long it = 0;
            for (;;) {
                if (it % steps_per_refill == 0) {
                    int live = 0;
                    for (auto& l : lane) live += l.active;
                    if (live < refill_below && ray_lo < ray_hi)
                        for (auto& l : lane) if (!l.active && ray_lo < ray_hi) {
                            l.active = true; l.fresh = true; l.r = make_ray(tris, ray_lo++, 777u);
                            l.G = NGroup{}; l.T.n = 0; l.st.clear();
                        }
                    live = 0;
                    for (auto& l : lane) live += l.active;
                    if (!live) break;
                }
                s.warp_iters += 1;
                for (auto& l : lane) s.lane_iters += l.active;
                for (int ns = 0; ns < pol.node_steps; ++ns) {
                // pop
                for (auto& l : lane) if (l.active && !l.fresh && !l.G.mask && !l.st.empty()) {
                    SEntry& e = l.st.back();
                    if (!e.is_tri) {
                        l.G = e.g; l.st.pop_back();
                        if (pol.cull_stale && l.G.bound > l.r.tbest) l.G.mask = 0;  // nothing in the group can be reached any more
                    }
                    else if (!l.T.n) { l.T = e.t; l.st.pop_back(); }
                }
                // node step
                int nl = 0;
                for (auto& l : lane) if (l.active && (l.fresh || l.G.mask)) {
                    ++nl;
                    int32_t visit;
                    if (l.fresh) { visit = 0; l.fresh = false; }
                    else {
                        const Wide::Node& pn = w.nodes[l.G.node];
                        int best = -1, bp = -1;
                        for (int c = 0; c < pn.n; ++c) if (l.G.mask >> c & 1) { const int pr = pn.slot[c] ^ l.r.oct; if (pr > bp) { bp = pr; best = c; } }
                        l.G.mask &= (uint8_t)~(1u << best);
                        if (l.G.mask) l.st.push_back(SEntry{false, l.G, TGroup{}});
                        visit = pn.child[best];
                    }
                    const Wide::Node& nd = w.nodes[visit];
                    NGroup ng; ng.node = visit;
                    TGroup nt;
                    float ctn[8];
                    for (int c = nd.n - 1; c >= 0; --c) {
                        float tn = l.r.tmin, tf = l.r.tbest;
                        for (int a = 0; a < 3; ++a) {
                            float t0 = (nd.cb[c].lo[a] - l.r.o[a]) * l.r.inv[a], t1 = (nd.cb[c].hi[a] - l.r.o[a]) * l.r.inv[a];
                            if (t0 > t1) std::swap(t0, t1);
                            tn = std::max(tn, t0); tf = std::min(tf, t1);
                        }
                        if (tn > tf) continue;
                        if (nd.child[c] >= 0) { ng.mask |= (uint8_t)(1u << c); ctn[c] = tn; }
                        else for (int k = 0; k < nd.ntri[c]; ++k) nt.tri[nt.n++] = nd.tri[c][k];
                    }
                    if (pol.cull_stale && ng.mask) {  // bound of the group that will be pushed: all members but the first visited
                        int first = -1, bp = -1;
                        for (int c = 0; c < nd.n; ++c) if (ng.mask >> c & 1) { const int pr = nd.slot[c] ^ l.r.oct; if (pr > bp) { bp = pr; first = c; } }
                        float m = FLT_MAX;
                        for (int c = 0; c < nd.n; ++c) if ((ng.mask >> c & 1) && c != first) m = std::min(m, ctn[c]);
                        ng.bound = m;
                    }
                    l.G = ng;
                    if (nt.n) {
                        if (l.T.n) l.st.push_back(SEntry{true, NGroup{}, l.T});
                        l.T = nt;
                    }
                }
                if (nl) { s.wnode += 1; s.nodes += nl; }
                }
                // triangle step
                int pending = 0; bool starving = false;
                for (auto& l : lane) if (l.active && l.T.n) { ++pending; if (!l.G.mask) starving = true; }
                if (pending && it % pol.tri_every == 0 && (pending >= pol.tri_min_lanes || starving)) {
                    s.wtri += 1;
                    for (auto& l : lane) if (l.active && l.T.n)
                        for (int q = 0; q < pol.tris_per_step && l.T.n; ++q) {
                            s.tris += 1;
                            const Tri& tq = tris[l.T.tri[--l.T.n]];
                            const V3 o{l.r.o[0], l.r.o[1], l.r.o[2]}, dir{l.r.d[0], l.r.d[1], l.r.d[2]};
                            const V3 e1 = tq.v[1] - tq.v[0], e2 = tq.v[2] - tq.v[0], p = cross(dir, e2);
                            const float det = dot(e1, p);
                            if (det == 0.f) continue;
                            const float id = 1.f / det; const V3 sv = o - tq.v[0];
                            const float uu = dot(sv, p) * id; const V3 qq = cross(sv, e1);
                            const float vv = dot(dir, qq) * id, tt = dot(e2, qq) * id;
                            if (uu >= 0.f && vv >= 0.f && uu + vv <= 1.f && tt >= l.r.tmin && tt < l.r.tbest) l.r.tbest = tt;
                        }
                }
                // terminate
                for (auto& l : lane) if (l.active && !l.fresh && !l.G.mask && !l.T.n && l.st.empty()) l.active = false;
                ++it;
            }
            part[wi] = s;
        }
    });
    for (auto& x : th) x.join();
    SimOut s;
    for (auto& p : part) { s.warp_iters += p.warp_iters; s.lane_iters += p.lane_iters; s.wnode += p.wnode; s.wtri += p.wtri; s.nodes += p.nodes; s.tris += p.tris; }
    s.cost = ((63.0 + 22.0 * (pol.node_steps - 1)) * s.warp_iters + (240.0 + (pol.cull_stale ? 16.0 : 0.0)) * s.wnode +
              74.0 * s.wtri * (1.0 + 0.6 * (pol.tris_per_step - 1))) / nrays;
    return s;
}

static void report_sim(const Wide& w, const std::vector<Tri>& tris, uint32_t nrays, const Policy& pol) {
    const SimOut s = simulate(w, tris, nrays, pol);
    printf("%-44s iters/ray %6.2f  nodes/ray %6.2f  tris/ray %5.2f  lanes: live %5.2f node %5.2f tri %5.2f  tri steps/iter %.3f  warp instr/ray %7.1f\n",
           pol.name, s.lane_iters / nrays, s.nodes / nrays, s.tris / nrays, s.lane_iters / s.warp_iters, s.nodes / std::max(s.wnode, 1.0),
           s.tris / std::max(s.wtri, 1.0), s.wtri / s.warp_iters, s.cost);
    fflush(stdout);
}

static void report(const char* name, Bin& b, const std::vector<Tri>& tris, uint32_t nrays, int quant = 0, int order = 0) {
    Wide w = collapse(b);
    assign_slots(w);
    if (quant) quantise(w, quant);
    const Stats s = trace_stats(w, tris, nrays, 12345u, order);
    printf("%-34s binary SAH %8.2f  BVH8 nodes %8zu  per ray: nodes %6.2f (no child hit %5.2f)  tris %5.2f  hit %.2f\n", name,
           sah_binary(b), w.nodes.size(), s.nodes, s.empty, s.tris, s.hits);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 1000000u;
    const uint32_t nrays = argc > 2 ? (uint32_t)atoi(argv[2]) : 400000u;
    const uint32_t seed = 0x5EED0002u;
    const float scale = (float)std::pow((double)n, -1.0 / 3.0);
    std::vector<Tri> tris(n);
    for (uint32_t i = 0; i < n; ++i) {  // shade.cu k_soup
        float f[12];
        for (uint32_t j = 0; j < 12; ++j) { uint32_t st = seed + (16u * i + j) * 0x9E3779B9u; f[j] = (float)pcg(st) * 2.3283064365386963e-10f; }
        const float c[3] = {f[0] * 2.f - 1.f, f[1] * 2.f - 2.f, f[2] * 2.f - 1.f};
        for (int v = 0; v < 3; ++v) {
            float p[3];
            for (int a = 0; a < 3; ++a) p[a] = c[a] + (f[3 + 3 * v + a] * 2.f - 1.f) * scale;
            tris[i].v[v] = {p[0], p[1], p[2]};
        }
    }
    printf("soup %u triangles, scale %.5f, %u rays\n", n, scale, nrays);
    {
        Bin b = build_lbvh(tris); refit(b);
        report("LBVH (what build.cu builds)", b, tris, nrays);
        report("  8-bit boxes, one exponent", b, tris, nrays, 1);
        report("  8-bit boxes, exponent per axis", b, tris, nrays, 2);
        report("  one exponent, stale visited", b, tris, nrays, 1, 1);
        report("  one exponent, octant order (GPU)", b, tris, nrays, 1, 2);
        report("  octant order + group-min culling", b, tris, nrays, 1, 3);
        report("  octant order + per-child culling", b, tris, nrays, 1, 4);
        if (argc > 3 && !strcmp(argv[3], "simt")) {
            Wide w = collapse(b);
            assign_slots(w);
            quantise(w, 1);
            printf("warp-level model of the traversal loop (32 lanes, refill below 30 every 2 iterations):\n");
            report_sim(w, tris, nrays, Policy{"k_trace: 1 triangle test per iteration"});
            report_sim(w, tris, nrays, Policy{"triangle step every 2nd iteration", 2, 1, 1});
            report_sim(w, tris, nrays, Policy{"2 triangle tests per step", 1, 1, 2});
            report_sim(w, tris, nrays, Policy{"all pending triangles per step", 1, 1, 16});
            report_sim(w, tris, nrays, Policy{"2 node steps per iteration", 1, 1, 1, 2});
            report_sim(w, tris, nrays, Policy{"3 node steps per iteration", 1, 1, 1, 3});
            report_sim(w, tris, nrays, Policy{"2 node steps, 2 triangle tests", 1, 1, 2, 2});
            report_sim(w, tris, nrays, Policy{"stale groups culled (+16 instr per node step)", 1, 1, 1, 1, true});
            report_sim(w, tris, nrays, Policy{"2 node steps + stale groups culled", 1, 1, 1, 2, true});
            for (int th : {8})
                { char nm[64]; snprintf(nm, sizeof nm, "step when >= %d lanes pending or one starves", th); report_sim(w, tris, nrays, Policy{nm, 1, th, 1}); }
            for (int th : {8})
                { char nm[64]; snprintf(nm, sizeof nm, ">= %d lanes or starving, 2 tests per step", th); report_sim(w, tris, nrays, Policy{nm, 1, th, 2}); }
            return 0;
        }
        for (int pass = 1; pass <= 3; ++pass) {
            const int r = rotate_pass(b); refit(b);
            char nm[64]; snprintf(nm, sizeof nm, "+ rotation pass %d (%d rotations)", pass, r);
            report(nm, b, tris, nrays);
        }
    }
    if (argc > 3 && !strcmp(argv[3], "split")) {
        for (int lv : {0, 1, 2, 3}) for (float ms : {0.f, 0.5f, 1.0f}) {
            if (lv == 0 && ms > 0.f) continue;
            std::vector<Ref> refs = presplit(tris, lv, ms * scale);
            Bin b = build_lbvh_refs(refs); refit(b);
            char nm[64]; snprintf(nm, sizeof nm, "split lv %d min %.1fs refs %.2fx", lv, ms, (double)refs.size() / n);
            report(nm, b, tris, nrays, 1, 2);
            rebuild_small_subtrees(b, 32);
            snprintf(nm, sizeof nm, "  + SAH<=32");
            report(nm, b, tris, nrays, 1, 2);
        }
        return 0;
    }
    for (uint32_t m : {32u, 512u, 0xffffffffu}) {
        Bin b = build_lbvh(tris); refit(b);
        rebuild_small_subtrees(b, m);
        char nm[64]; snprintf(nm, sizeof nm, "SAH rebuild of subtrees <= %u", m);
        report(nm, b, tris, nrays);
    }
    return 0;
}
