// build.cu — acceleration-structure build on the GPU.
//
// Replaces the driver work behind Accel::Accel -> buildAccelerationStructuresKHR
// (reference main.cpp:416-450; BLAS at :512, TLAS at :538) with:
//   K1  prim_bounds      per-primitive AABB + scene bounds (warp-shuffle reduce, ordered-int atomics)
//   K2  morton_keys      30-bit Morton code of the AABB centre; 64-bit key = morton << 32 | prim
//   K3  radix sort       radix_sort.cu (hand-written 8-bit LSD onesweep-style passes)
//   K4  lbvh_hierarchy   Karras 2012: one thread per internal node, clz on key XOR
//   K5  lbvh_refit       bottom-up AABBs with per-node arrival counters
//   K6  bvh8_collapse    binary -> 8-wide, greedy largest-area opening, octant slot assignment,
//                        quantisation to the 64-byte Node8 (level-synchronous; nodes and triangle
//                        records share one array, a node's children records are contiguous)
//   K7  woop_transform   64-byte records (unit-triangle transform + primitive id) at the record
//                        positions the collapse reserved
// Everything is generic over "primitives with an AABB" so the same code builds the instance-level
// BVH8 of two-level scenes.
#include <cfloat>
#include <cstdio>

#include "build.cuh"

namespace {

constexpr int kBlock = 256;
inline unsigned grid_for(uint64_t n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }

// ---------------------------------------------------------------- ordered-int float atomics
__device__ __forceinline__ uint32_t enc_f(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_f(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

__global__ void k_init_bounds(uint32_t* b) {
    if (threadIdx.x < 3) b[threadIdx.x] = 0xffffffffu;       // min
    else if (threadIdx.x < 6) b[threadIdx.x] = 0u;           // max
}

__device__ __forceinline__ void warp_reduce_bounds(float lo[3], float hi[3], uint32_t* bounds) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bounds[a], enc_f(lo[a]));
            atomicMax(&bounds[3 + a], enc_f(hi[a]));
        }
    }
}

// K1 (triangles): AABB of triangle i from the indexed vertex buffer (closesthit.rchit:52-54 layout). A triangle with a
// NaN or infinite vertex is INACTIVE, as in the Vulkan acceleration-structure build the reference calls (a NaN in the
// first vertex component marks an inactive triangle there): it gets the empty box (lo = +max, hi = -max), which leaves
// the scene bounds and every enclosing node box alone, quantises to an empty slab in the BVH8 and is never hit.
__global__ void k_tri_bounds(const float* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t n,
                             float4* __restrict__ plo, float4* __restrict__ phi, uint32_t* bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        bool finite = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* v = verts + 3 * (size_t)idx[3 * (size_t)i + c];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                finite = finite && isfinite(v[a]);
                lo[a] = fminf(lo[a], v[a]);
                hi[a] = fmaxf(hi[a], v[a]);
            }
        }
        if (!finite) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
        }
        plo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        phi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    warp_reduce_bounds(lo, hi, bounds);
}

// K1 (instances): world AABB of instance i = transformed corners of the mesh bounds.
__global__ void k_instance_bounds(const float* __restrict__ xf, uint32_t n, float3 mlo, float3 mhi,
                                  float4* __restrict__ plo, float4* __restrict__ phi, uint32_t* bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        const float* m = xf + 12 * (size_t)i;
        for (int c = 0; c < 8; ++c) {
            float x = (c & 1) ? mhi.x : mlo.x, y = (c & 2) ? mhi.y : mlo.y, z = (c & 4) ? mhi.z : mlo.z;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float w = m[4 * r] * x + m[4 * r + 1] * y + m[4 * r + 2] * z + m[4 * r + 3];
                // widen by a few ulp: the shade kernel transforms vertices without FMA
                float e = 4e-7f * (fabsf(m[4 * r] * x) + fabsf(m[4 * r + 1] * y) + fabsf(m[4 * r + 2] * z) + fabsf(m[4 * r + 3]));
                lo[r] = fminf(lo[r], w - e);
                hi[r] = fmaxf(hi[r], w + e);
            }
        }
        plo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        phi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    warp_reduce_bounds(lo, hi, bounds);
}

// K2: 10 bits per axis, x most significant inside each triple.
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton30(float x, float y, float z) {
    uint32_t qx = (uint32_t)fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    uint32_t qy = (uint32_t)fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
    uint32_t qz = (uint32_t)fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
    return expand10(qx) * 4u + expand10(qy) * 2u + expand10(qz);
}
__global__ void k_morton_keys(const float4* __restrict__ plo, const float4* __restrict__ phi, uint32_t n,
                              const uint32_t* __restrict__ bounds, uint64_t* __restrict__ keys) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 blo = make_float3(dec_f(bounds[0]), dec_f(bounds[1]), dec_f(bounds[2]));
    float3 bhi = make_float3(dec_f(bounds[3]), dec_f(bounds[4]), dec_f(bounds[5]));
    float4 lo = plo[i], hi = phi[i];
    float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
    float ex = bhi.x - blo.x, ey = bhi.y - blo.y, ez = bhi.z - blo.z;
    float nx = ex > 0.f ? (cx - blo.x) / ex : 0.f;
    float ny = ey > 0.f ? (cy - blo.y) / ey : 0.f;
    float nz = ez > 0.f ? (cz - blo.z) / ez : 0.f;
    keys[i] = ((uint64_t)morton30(nx, ny, nz) << 32) | i;
}

// K4: Karras 2012. Keys are unique (primitive id in the low word), so delta needs no tie-break.
// Node ids: internal i in [0,n-1); leaf k is n-1+k.
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, uint64_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(ki ^ keys[j]));
}
__global__ void k_lbvh_hierarchy(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ left,
                                 uint32_t* __restrict__ right, uint32_t* __restrict__ parent,
                                 uint32_t* __restrict__ first, uint32_t* __restrict__ last) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int ni = (int)n;
    if (i >= ni - 1) return;
    uint64_t ki = keys[i];
    int d = (delta(keys, ni, ki, i + 1) - delta(keys, ni, ki, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, ni, ki, i - d);
    int lmax = 2;
    while (delta(keys, ni, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, ni, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, ni, ki, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, ni, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    uint32_t lc = (lo == gamma) ? (uint32_t)(ni - 1 + gamma) : (uint32_t)gamma;
    uint32_t rc = (hi == gamma + 1) ? (uint32_t)(ni - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc] = (uint32_t)i;
    parent[rc] = (uint32_t)i;
    first[i] = (uint32_t)lo;
    last[i] = (uint32_t)hi;
    if (i == 0) parent[0] = 0xffffffffu;
}

// K5: leaves copy their primitive box, then climb; the second arrival at a node merges.
// Triangles per leaf slot. The record format allows 3; measured on the 10 M soup: 3 -> 28.2 nodes + 8.6 triangles per
// ray, 2 -> 29.3 + 5.5 and +2.4 % Mray/s, 1 -> 29.9 + 4.7 and the same speed with 17 % more nodes.
constexpr uint32_t kMaxLeafTris = 2u;
__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
    // the empty box of an inactive primitive (hi < lo) has no area
    float dx = fmaxf(hi.x - lo.x, 0.f), dy = fmaxf(hi.y - lo.y, 0.f), dz = fmaxf(hi.z - lo.z, 0.f);
    return dx * dy + dy * dz + dz * dx;
}

// SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3), the bottom-up half, fused into the refit:
//   c(n,1) = min(c_leaf(n), c_dist(n,8) + A_n * c_node),  c(n,i) = min(c_dist(n,i), c(n,i-1)) for i = 2..7,
//   c_dist(n,j) = min over 0<k<j of c(left,k) + c(right,j-k),  c_leaf(n) = A_n * P_n * c_prim if P_n <= kMaxLeafTris
// c(n,i) is the cheapest way to hand the subtree of n to a parent that offers it i child slots. dp_cost holds
// c(n,1..7) per binary node, dp_dec the choices: byte 0 = 1 if n ends as a leaf slot; byte j-1 (j = 2..8) = the k of
// c_dist(n,j), with 0x80 set (j <= 7) when c(n,j) falls back to c(n,j-1).
constexpr float kCostNode = 1.0f, kCostPrim = 0.6f;
__device__ __forceinline__ void dp_leaf(float* __restrict__ cost, uint8_t* __restrict__ dec, uint32_t node, float area) {
#pragma unroll
    for (int i = 0; i < 7; ++i) cost[7 * (size_t)node + i] = area * kCostPrim;
    dec[8 * (size_t)node] = 1;
#pragma unroll
    for (int j = 1; j < 8; ++j) dec[8 * (size_t)node + j] = 0x80;
}
__device__ __forceinline__ void dp_internal(float* __restrict__ cost, uint8_t* __restrict__ dec, uint32_t node, uint32_t l,
                                            uint32_t r, float area, uint32_t count) {
    float cl[7], cr[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) { cl[i] = __ldcg(&cost[7 * (size_t)l + i]); cr[i] = __ldcg(&cost[7 * (size_t)r + i]); }
    float cd[9];
    uint8_t kb[9];
#pragma unroll
    for (int j = 2; j <= 8; ++j) {
        float best = FLT_MAX;
        int bk = 1;
#pragma unroll
        for (int k = 1; k < j; ++k) {
            if (k > 7 || j - k > 7) continue;
            const float c = cl[k - 1] + cr[j - k - 1];
            // equal costs (coincident primitives): the more balanced split, or the collapse degenerates into a chain
            if (c < best || (c == best && abs(2 * k - j) < abs(2 * bk - j))) { best = c; bk = k; }
        }
        cd[j] = best; kb[j] = (uint8_t)bk;
    }
    float c[8];
    const bool leaf = count <= kMaxLeafTris;
    c[1] = leaf ? area * (float)count * kCostPrim : cd[8] + area * kCostNode;
    uint8_t d[8];
    d[0] = leaf ? 1 : 0;
    d[7] = kb[8];
#pragma unroll
    for (int i = 2; i <= 7; ++i) {
        if (cd[i] < c[i - 1]) { c[i] = cd[i]; d[i - 1] = kb[i]; }
        else { c[i] = c[i - 1]; d[i - 1] = (uint8_t)(kb[i] | 0x80); }
    }
#pragma unroll
    for (int i = 1; i <= 7; ++i) cost[7 * (size_t)node + i - 1] = c[i];
#pragma unroll
    for (int j = 0; j < 8; ++j) dec[8 * (size_t)node + j] = d[j];
}

// K5: leaves copy their primitive box, then climb; the second arrival at a node merges (boxes and collapse costs).
__global__ void k_lbvh_refit(const uint32_t* __restrict__ leaf_prim, uint32_t n, const float4* __restrict__ plo,
                             const float4* __restrict__ phi, const uint32_t* __restrict__ left,
                             const uint32_t* __restrict__ right, const uint32_t* __restrict__ parent,
                             const uint32_t* __restrict__ first, const uint32_t* __restrict__ last,
                             uint32_t* __restrict__ arrive, float4* __restrict__ nlo, float4* __restrict__ nhi,
                             float* __restrict__ dp_cost, uint8_t* __restrict__ dp_dec) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t prim = leaf_prim[k];
    uint32_t node = n - 1 + k;
    float4 lo = plo[prim], hi = phi[prim];
    nlo[node] = lo;
    nhi[node] = hi;
    dp_leaf(dp_cost, dp_dec, node, half_area(lo, hi));
    if (n == 1) return;
    uint32_t p = parent[node];
    while (p != 0xffffffffu) {
        __threadfence();
        if (atomicAdd(&arrive[p], 1u) == 0u) return;  // first to arrive: sibling not ready
        uint32_t a = left[p], b = right[p];
        float4 alo = __ldcg(&nlo[a]), ahi = __ldcg(&nhi[a]), blo = __ldcg(&nlo[b]), bhi = __ldcg(&nhi[b]);
        lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
        hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
        nlo[p] = lo;
        nhi[p] = hi;
        dp_internal(dp_cost, dp_dec, p, a, b, half_area(lo, hi), last[p] - first[p] + 1u);
        p = parent[p];
    }
}

// ---------------------------------------------------------------- K6: collapse to BVH8
struct CollapseArgs {
    uint32_t n;  // primitives
    const uint32_t* leaf_prim;  // leaf position -> primitive
    const uint32_t *left, *right, *first, *last;
    const float4 *nlo, *nhi;
    Node8* recs;              // the record array; this kernel writes the node records
    uint32_t* wide_src;       // record index of a node -> binary node id it expands
    uint32_t* rec_prim;       // record index -> primitive id (triangle records; BPT_MISS for nodes)
    uint32_t* counters;       // [0] records allocated, [1] nodes queued for the next level, [2] triangle records
    const uint32_t* list_cur; // records of this level's nodes
    uint32_t* list_next;      // records of the next level's nodes
    uint32_t level_count;
    float pad;                // conservative widening of every quantised box (world units)
    const uint8_t* dp_dec;    // choices of the optimal collapse (k_lbvh_refit), or null: greedy area-first opening
    float gbias[3], gstep[3]; // scene grid the node origins are quantised on (BPT_GRID_BITS per axis), build.cuh
};

__device__ __forceinline__ uint32_t node_count(const CollapseArgs& a, uint32_t node) {
    return node >= a.n - 1 ? 1u : a.last[node] - a.first[node] + 1u;
}
__device__ __forceinline__ uint32_t node_first(const CollapseArgs& a, uint32_t node) {
    return node >= a.n - 1 ? node - (a.n - 1) : a.first[node];
}

constexpr float kQuantMargin = 0.0078125f;  // 2^-7 of a quantisation step, see k_bvh8_collapse

// exponent e with 2^e * 255 >= ext (with margin), clamped away from denormals
__device__ __forceinline__ int quant_exponent(float ext) {
    if (!(ext > 0.f)) return -100;
    int k;
    frexpf(ext * (1.02f / 255.0f), &k);  // value = m * 2^k, m in [0.5,1)  =>  2^k > value; 2 % = 5 steps of slack
    return max(-100, min(k, 120));
}

__global__ void k_bvh8_collapse(CollapseArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.level_count) return;
    const uint32_t w = a.list_cur[li];
    const uint32_t src = a.wide_src[w];

    uint32_t cand[8];
    float area[8];
    uint32_t cnt[8];
    int nc = 1;
    if (a.dp_dec && a.n > 1 && src < a.n - 1) {
        // top-down half of the optimal collapse: the children of this wide node are what c_dist(src, 8) chose
        nc = 0;
        uint32_t st_node[8];
        int st_slots[8];
        int top = 0;
        const uint32_t k8 = a.dp_dec[8 * (size_t)src + 7] & 0x7fu;
        st_node[top] = a.right[src]; st_slots[top++] = 8 - (int)k8;
        st_node[top] = a.left[src]; st_slots[top++] = (int)k8;
        while (top) {
            const uint32_t nd = st_node[--top];
            int i = st_slots[top];
            const uint8_t* d = a.dp_dec + 8 * (size_t)nd;
            while (i > 1 && (d[i - 1] & 0x80u)) --i;  // c(nd,i) == c(nd,i-1)
            if (i == 1 || nd >= a.n - 1) {
                cand[nc] = nd; cnt[nc] = node_count(a, nd); area[nc] = 0.f; ++nc;
            } else {
                const int k = d[i - 1] & 0x7f;
                st_node[top] = a.right[nd]; st_slots[top++] = i - k;
                st_node[top] = a.left[nd]; st_slots[top++] = k;
            }
        }
    } else {
    cand[0] = src;
    cnt[0] = node_count(a, src);
    area[0] = FLT_MAX;
    // greedy: open the largest-area candidate that still holds >= 2 primitives
    while (nc < 8) {
        int best = -1;
        float barea = -1.f;
        for (int c = 0; c < nc; ++c)
            if (cnt[c] >= 2u && area[c] > barea) { barea = area[c]; best = c; }
        if (best < 0) break;
        uint32_t b = cand[best];
        uint32_t l = a.left[b], r = a.right[b];
        cand[best] = l;
        cnt[best] = node_count(a, l);
        area[best] = half_area(a.nlo[l], a.nhi[l]);
        cand[nc] = r;
        cnt[nc] = node_count(a, r);
        area[nc] = half_area(a.nlo[r], a.nhi[r]);
        ++nc;
    }
    }

    const float4 blo = a.nlo[src], bhi = a.nhi[src];
    const float pad = a.pad;
    // Child-box grid: ONE exponent for the three axes, from the largest padded extent plus the resolution of the
    // origin grid (the origin is rounded down onto it) with ~5 steps of slack; the origin sits one margin below the
    // padded box, on the scene grid.
    const float ext = fmaxf(fmaxf((bhi.x - blo.x) + a.gstep[0], (bhi.y - blo.y) + a.gstep[1]), (bhi.z - blo.z) + a.gstep[2]);
    const int ex = quant_exponent(ext + 2.f * pad);
    const float sx = exp2f((float)ex), isx = exp2f((float)-ex);
    const float sy = sx, sz = sx, isy = isx, isz = isx;
    uint32_t org[3];
    float porg[3];
    {
        const float target[3] = {blo.x - pad - kQuantMargin * sx, blo.y - pad - kQuantMargin * sy, blo.z - pad - kQuantMargin * sz};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // decode(c) is what the traversal kernel evaluates: fmaf(float(2^23 + c), gstep, gbias)
            auto decode = [&](int c) { return fmaf(8388608.0f + (float)c, a.gstep[k], a.gbias[k]); };
            int c = (int)floorf((target[k] - decode(0)) / a.gstep[k]);
            c = max(0, min(c, (1 << BPT_GRID_BITS) - 1));
            while (c > 0 && decode(c) > target[k]) --c;
            org[k] = (uint32_t)c;
            porg[k] = decode(c);
        }
    }
    const float px = porg[0], py = porg[1], pz = porg[2];
    const float cx = 0.5f * (blo.x + bhi.x), cy = 0.5f * (blo.y + bhi.y), cz = 0.5f * (blo.z + bhi.z);

    // slot assignment: slot bit 2/1/0 set = child lies on the +x/+y/+z side of the node centre.
    // Greedy maximum of dot(sign(slot), child centre - node centre) over unassigned pairs.
    float4 clo[8], chi[8];
    float ddx[8], ddy[8], ddz[8];
    for (int c = 0; c < nc; ++c) {
        clo[c] = a.nlo[cand[c]];
        chi[c] = a.nhi[cand[c]];
        ddx[c] = 0.5f * (clo[c].x + chi[c].x) - cx;
        ddy[c] = 0.5f * (clo[c].y + chi[c].y) - cy;
        ddz[c] = 0.5f * (clo[c].z + chi[c].z) - cz;
    }
    int slot_of[8];
    int child_in[8];
    for (int s = 0; s < 8; ++s) child_in[s] = -1;
    for (int c = 0; c < nc; ++c) slot_of[c] = -1;
    for (int it = 0; it < nc; ++it) {
        float bs = -FLT_MAX;
        int bc = -1, bsl = -1;
        for (int c = 0; c < nc; ++c) {
            if (slot_of[c] >= 0) continue;
            for (int s = 0; s < 8; ++s) {
                if (child_in[s] >= 0) continue;
                float sc = ((s & 4) ? ddx[c] : -ddx[c]) + ((s & 2) ? ddy[c] : -ddy[c]) + ((s & 1) ? ddz[c] : -ddz[c]);
                if (sc > bs) { bs = sc; bc = c; bsl = s; }
            }
        }
        slot_of[bc] = bsl;
        child_in[bsl] = bc;
    }

    // allocate the children records: internal children first (slot order), then the leaf triangles
    uint32_t n_int = 0, n_leaf = 0;
    for (int c = 0; c < nc; ++c) {
        if (cnt[c] > kMaxLeafTris) ++n_int; else n_leaf += cnt[c];
    }
    const uint32_t child_base = atomicAdd(&a.counters[0], n_int + n_leaf);
    const uint32_t qbase = n_int ? atomicAdd(&a.counters[1], n_int) : 0u;
    if (n_leaf) atomicAdd(&a.counters[2], n_leaf);
    const uint32_t tri_base = child_base + n_int;

    Node8 nd;
    nd.org_lo = org[0] | (org[1] << 21);
    nd.org_hi = (org[1] >> 11) | (org[2] << 10);
    nd.child_base = child_base;
    uint32_t valid = 0, int_rank = 0, leaf_off = 0;
    for (int s = 0; s < 8; ++s) {
        int c = child_in[s];
        if (c < 0) {
            nd.qlox[s] = nd.qloy[s] = nd.qloz[s] = 255;
            nd.qhix[s] = nd.qhiy[s] = nd.qhiz[s] = 0;
            continue;
        }
        // conservative quantisation with margin m = kQuantMargin steps:
        //   origin + qlo*scale <= lo - pad - m*scale,  origin + qhi*scale >= hi + pad + m*scale
        // (the traversal kernel evaluates the planes with an error below 2^-9 step, trace.cu byte_f)
        auto qlo = [&](float v, float p, float sc, float isc) {
            float t = v - pad - kQuantMargin * sc;
            int q = (int)floorf((t - p) * isc);
            q = max(0, min(q, 255));
            while (q > 0 && p + (float)q * sc > t) --q;
            return (uint8_t)q;
        };
        auto qhi = [&](float v, float p, float sc, float isc) {
            float t = v + pad + kQuantMargin * sc;
            int q = (int)ceilf((t - p) * isc);
            q = max(0, min(q, 255));
            while (q < 255 && p + (float)q * sc < t) ++q;
            return (uint8_t)q;
        };
        nd.qlox[s] = qlo(clo[c].x, px, sx, isx); nd.qhix[s] = qhi(chi[c].x, px, sx, isx);
        nd.qloy[s] = qlo(clo[c].y, py, sy, isy); nd.qhiy[s] = qhi(chi[c].y, py, sy, isy);
        nd.qloz[s] = qlo(clo[c].z, pz, sz, isz); nd.qhiz[s] = qhi(chi[c].z, pz, sz, isz);
        if (cnt[c] > kMaxLeafTris) {
            valid |= 1u << (16 + s);
            a.wide_src[child_base + int_rank] = cand[c];
            a.list_next[qbase + int_rank] = child_base + int_rank;
            ++int_rank;
        } else {
            uint32_t k = cnt[c];
            valid |= k << (2 * s);
            uint32_t f = node_first(a, cand[c]);
            for (uint32_t j = 0; j < k; ++j)
                a.rec_prim[tri_base + leaf_off + j] = a.leaf_prim[f + j];
            leaf_off += k;
        }
    }
    nd.e_valid = valid | ((uint32_t)(ex + 127) << 24);
    uint4* dst = reinterpret_cast<uint4*>(a.recs + w);
    const uint4* srcw = reinterpret_cast<const uint4*>(&nd);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = srcw[q];
}

// K7: Woop transform of the triangle a record holds. Rows are computed in double and rounded once.
__global__ void k_woop(const float* __restrict__ verts, const uint32_t* __restrict__ idx,
                       const uint32_t* __restrict__ rec_prim, uint32_t nrecs, Node8* __restrict__ recs) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nrecs) return;
    uint32_t prim = rec_prim[s];
    if (prim == BPT_MISS) return;  // a node record
    WoopTri* out = reinterpret_cast<WoopTri*>(recs);
    double v[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* p = verts + 3 * (size_t)idx[3 * (size_t)prim + c];
        v[c][0] = p[0]; v[c][1] = p[1]; v[c][2] = p[2];
    }
    double e1[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
    double e2[3] = {v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2]};
    double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    double det = nx * nx + ny * ny + nz * nz;  // det [e1 e2 n]
    WoopTri w;
    w.prim = prim; w.pad0 = w.pad1 = w.pad2 = 0;
    if (!(det > 0.0) || !isfinite(det)) {
        w.ru = w.rv = w.rw = make_float4(0.f, 0.f, 0.f, 0.f);  // degenerate: d'.z == 0 -> never hit
    } else {
        double inv = 1.0 / det;
        // rows of [e1 e2 n]^-1: (e2 x n)/det, (n x e1)/det, n/det
        double ru[3] = {(e2[1] * nz - e2[2] * ny) * inv, (e2[2] * nx - e2[0] * nz) * inv, (e2[0] * ny - e2[1] * nx) * inv};
        double rv[3] = {(ny * e1[2] - nz * e1[1]) * inv, (nz * e1[0] - nx * e1[2]) * inv, (nx * e1[1] - ny * e1[0]) * inv};
        double rw[3] = {nx * inv, ny * inv, nz * inv};
        double cu = -(ru[0] * v[0][0] + ru[1] * v[0][1] + ru[2] * v[0][2]);
        double cv = -(rv[0] * v[0][0] + rv[1] * v[0][1] + rv[2] * v[0][2]);
        double cw = -(rw[0] * v[0][0] + rw[1] * v[0][1] + rw[2] * v[0][2]);
        w.ru = make_float4((float)ru[0], (float)ru[1], (float)ru[2], (float)cu);
        w.rv = make_float4((float)rv[0], (float)rv[1], (float)rv[2], (float)cv);
        w.rw = make_float4((float)rw[0], (float)rw[1], (float)rw[2], (float)cw);
    }
    out[s] = w;
}

// K8: instance records of the instance-level BVH8: rows of the inverse 3x4 transform + instance id
__global__ void k_instance_records(const float* __restrict__ inv, const uint32_t* __restrict__ rec_prim, uint32_t nrecs,
                                   Node8* __restrict__ recs) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nrecs) return;
    const uint32_t inst = rec_prim[s];
    if (inst == BPT_MISS) return;
    const float* m = inv + 12 * (size_t)inst;
    WoopTri w;
    w.ru = make_float4(m[0], m[1], m[2], m[3]);
    w.rv = make_float4(m[4], m[5], m[6], m[7]);
    w.rw = make_float4(m[8], m[9], m[10], m[11]);
    w.prim = inst; w.pad0 = w.pad1 = w.pad2 = 0;
    reinterpret_cast<WoopTri*>(recs)[s] = w;
}

// K8: copies the instance-level records behind the mesh records, rebasing the child pointers of its nodes
__global__ void k_append_recs(const Node8* __restrict__ src, const uint32_t* __restrict__ rec_prim, uint32_t n,
                              uint32_t rec_off, Node8* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Node8 nd = src[i];
    if (rec_prim[i] == BPT_MISS) nd.child_base += rec_off;
    dst[i] = nd;
}

template <class T>
cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 16);
}

}  // namespace

void bvh8_free(Bvh8& b) {
    cudaFree(b.plo); cudaFree(b.phi); cudaFree(b.keys); cudaFree(b.keys_tmp); cudaFree(b.left); cudaFree(b.right);
    cudaFree(b.parent); cudaFree(b.first); cudaFree(b.last); cudaFree(b.arrive); cudaFree(b.nlo); cudaFree(b.nhi);
    cudaFree(b.dp_cost); cudaFree(b.dp_dec); cudaFree(b.leaf_prim);
    cudaFree(b.recs); cudaFree(b.wide_src); cudaFree(b.rec_prim); cudaFree(b.list[0]); cudaFree(b.list[1]);
    cudaFree(b.counters); cudaFree(b.bounds); cudaFree(b.sort_tmp);
    b = Bvh8{};
}

cudaError_t bvh8_alloc(Bvh8& b, uint32_t n) {
    if (b.n == n && b.recs) return cudaSuccess;  // rebuild over the same primitive count: keep the buffers
    const bool keep_collapse = b.optimal_collapse;
    const uint32_t keep_sah = b.sah_max_leaves;
    bvh8_free(b);
    b.n = n;
    b.optimal_collapse = keep_collapse;
    b.sah_max_leaves = keep_sah;
    cudaError_t e;
#define A(x) if ((e = (x)) != cudaSuccess) return e
    A(dalloc(&b.plo, n)); A(dalloc(&b.phi, n));
    A(dalloc(&b.keys, n)); A(dalloc(&b.keys_tmp, n));
    A(dalloc(&b.left, n)); A(dalloc(&b.right, n)); A(dalloc(&b.first, n)); A(dalloc(&b.last, n));
    A(dalloc(&b.parent, 2 * (size_t)n)); A(dalloc(&b.arrive, n));
    A(dalloc(&b.nlo, 2 * (size_t)n)); A(dalloc(&b.nhi, 2 * (size_t)n));
    A(dalloc(&b.dp_cost, 14 * (size_t)n)); A(dalloc(&b.dp_dec, 16 * (size_t)n));
    A(dalloc(&b.leaf_prim, n));
    // every wide node expands at least one binary internal node, so n nodes always suffice; + n triangle records
    b.nodes_cap = n < 8 ? 8 : n;
    b.recs_cap = b.nodes_cap + n;
    A(dalloc(&b.recs, b.recs_cap)); A(dalloc(&b.wide_src, b.recs_cap)); A(dalloc(&b.rec_prim, b.recs_cap));
    A(dalloc(&b.list[0], b.nodes_cap)); A(dalloc(&b.list[1], b.nodes_cap));
    A(dalloc(&b.counters, 8)); A(dalloc(&b.bounds, 8));
    b.sort_tmp_bytes = radix_sort_u64_temp_bytes(n);
    A(cudaMalloc(&b.sort_tmp, b.sort_tmp_bytes));
#undef A
    return cudaSuccess;
}

void bvh8_launch_tri_bounds(Bvh8& b, const float* verts, const uint32_t* idx, cudaStream_t st) {
    k_init_bounds<<<1, 32, 0, st>>>(b.bounds);
    k_tri_bounds<<<grid_for(b.n), kBlock, 0, st>>>(verts, idx, b.n, b.plo, b.phi, b.bounds);
}
void bvh8_launch_instance_bounds(Bvh8& b, const float* xforms, const float mesh_lo[3], const float mesh_hi[3],
                                 cudaStream_t st) {
    k_init_bounds<<<1, 32, 0, st>>>(b.bounds);
    k_instance_bounds<<<grid_for(b.n), kBlock, 0, st>>>(xforms, b.n, make_float3(mesh_lo[0], mesh_lo[1], mesh_lo[2]),
                                                         make_float3(mesh_hi[0], mesh_hi[1], mesh_hi[2]), b.plo,
                                                         b.phi, b.bounds);
}

// Runs K2..K6 on the primitive boxes already in b.plo/b.phi/b.bounds. Synchronises the stream
// once per BVH8 level (the reference's build is synchronous as well: main.cpp:236-237).
cudaError_t bvh8_build(Bvh8& b, cudaStream_t st) {
    const uint32_t n = b.n;
    cudaError_t e;
    k_morton_keys<<<grid_for(n), kBlock, 0, st>>>(b.plo, b.phi, n, b.bounds, b.keys);
    // K3: sort the 30 Morton bits [32,62) only (4 digit passes): the primitive id in the low word starts out ascending
    // and the sort is stable, so equal Morton codes stay ordered by primitive id (the duplicate tie-break, T9)
    uint64_t* sorted = radix_sort_u64(b.keys, b.keys_tmp, n, 32, 62, b.sort_tmp, b.sort_tmp_bytes, st);
    if (sorted != b.keys) { uint64_t* t = b.keys; b.keys = b.keys_tmp; b.keys_tmp = t; }
    if ((e = cudaMemsetAsync(b.arrive, 0, sizeof(uint32_t) * n, st)) != cudaSuccess) return e;
    if (n > 1) k_lbvh_hierarchy<<<grid_for(n - 1), kBlock, 0, st>>>(b.keys, n, b.left, b.right, b.parent, b.first, b.last);
    sah_launch_leaf_order(b.keys, n, b.leaf_prim, st);
    // quality stage: SAH rebuild of the subtrees with at most sah_max_leaves leaves (scratch: the level list and a spare
    // counter, both unused until the collapse)
    if (b.sah_max_leaves >= 3)
        sah_launch_rebuild(n, b.sah_max_leaves, b.plo, b.phi, b.leaf_prim, b.left, b.right, b.parent, b.first, b.last,
                           b.list[0], b.counters + 4, st);
    k_lbvh_refit<<<grid_for(n), kBlock, 0, st>>>(b.leaf_prim, n, b.plo, b.phi, b.left, b.right, b.parent, b.first, b.last, b.arrive,
                                                 b.nlo, b.nhi, b.dp_cost, b.dp_dec);

    uint32_t hb[6];
    if ((e = cudaMemcpyAsync(hb, b.bounds, sizeof(hb), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    float diag = 0.f, mag = 0.f;
    for (int a = 0; a < 3; ++a) {
        b.scene_lo[a] = dec_f(hb[a]);
        b.scene_hi[a] = dec_f(hb[3 + a]);
        float d = b.scene_hi[a] - b.scene_lo[a];
        diag += d * d;
        mag = fmaxf(mag, fmaxf(fabsf(b.scene_lo[a]), fabsf(b.scene_hi[a])));
    }
    diag = sqrtf(diag);

    // scene grid of the node origins: BPT_GRID_BITS per axis over the padded scene box. A node origin sits one
    // quantisation margin (2^-7 of the node's step, and a node has ONE step for its three axes: at most 1/127 of the
    // scene's LARGEST extent) below its box, so the grid reaches that far below the scene on every axis — also on an
    // axis along which the scene is flat (a single quad), where the margin is far larger than the extent itself.
    const float pad = 9.5367431640625e-07f * fmaxf(diag, mag);  // 2^-20 of the scene scale
    float smax = 0.f;
    for (int a = 0; a < 3; ++a) smax = fmaxf(smax, b.scene_hi[a] - b.scene_lo[a]);
    for (int a = 0; a < 3; ++a) {
        const float below = 2.f * pad + ((b.scene_hi[a] - b.scene_lo[a]) + smax) * 0.001f;
        const float ext = (b.scene_hi[a] - b.scene_lo[a]) + 2.f * pad + below + 1e-30f;
        b.grid_lo[a] = b.scene_lo[a] - below;
        b.grid_step[a] = ext * 1.002f / (float)(1 << BPT_GRID_BITS);
        b.grid_bias[a] = b.grid_lo[a] - 8388608.0f * b.grid_step[a];
    }

    CollapseArgs a;
    a.n = n; a.leaf_prim = b.leaf_prim; a.left = b.left; a.right = b.right; a.first = b.first; a.last = b.last;
    a.nlo = b.nlo; a.nhi = b.nhi; a.recs = b.recs; a.wide_src = b.wide_src; a.rec_prim = b.rec_prim;
    a.counters = b.counters;
    a.dp_dec = b.optimal_collapse ? b.dp_dec : nullptr;
    a.pad = pad;
    for (int k = 0; k < 3; ++k) { a.gbias[k] = b.grid_bias[k]; a.gstep[k] = b.grid_step[k]; }
    const uint32_t init[3] = {1u, 0u, 0u};  // record 0 = root node, expands the binary root (internal 0, or leaf 0 when n == 1)
    const uint32_t zero = 0u;
    if ((e = cudaMemsetAsync(b.rec_prim, 0xff, sizeof(uint32_t) * b.recs_cap, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(b.counters, init, sizeof(init), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(b.wide_src, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(b.list[0], &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    uint32_t count = 1, depth = 0, nodes = 0;
    int cur = 0;
    while (count) {
        a.list_cur = b.list[cur];
        a.list_next = b.list[cur ^ 1];
        a.level_count = count;
        k_bvh8_collapse<<<grid_for(count, 64), 64, 0, st>>>(a);
        uint32_t hc[3];
        if ((e = cudaMemcpyAsync(hc, b.counters, sizeof(hc), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(b.counters + 1, &zero, sizeof(uint32_t), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++depth;
        nodes += count;
        count = hc[1];
        b.num_recs = hc[0];
        b.num_leaf_slots = hc[2];
        if (b.num_recs > b.recs_cap || nodes + count > b.nodes_cap) return cudaErrorMemoryAllocation;
        cur ^= 1;
    }
    b.num_nodes = nodes;
    b.depth = depth;
    return cudaSuccess;
}

void bvh8_launch_woop(const Bvh8& b, const float* verts, const uint32_t* idx, cudaStream_t st) {
    k_woop<<<grid_for(b.num_recs), kBlock, 0, st>>>(verts, idx, b.rec_prim, b.num_recs, b.recs);
}

void bvh8_launch_instance_records(const Bvh8& tlas, const float* inv_xforms, cudaStream_t st) {
    k_instance_records<<<grid_for(tlas.num_recs), kBlock, 0, st>>>(inv_xforms, tlas.rec_prim, tlas.num_recs, tlas.recs);
}
void bvh8_launch_append_recs(const Bvh8& tlas, uint32_t rec_off, Node8* dst, cudaStream_t st) {
    k_append_recs<<<grid_for(tlas.num_recs), kBlock, 0, st>>>(tlas.recs, tlas.rec_prim, tlas.num_recs, rec_off, dst);
}
