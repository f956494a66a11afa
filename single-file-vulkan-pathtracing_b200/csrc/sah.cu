// sah.cu — BVH quality stage between K4 (Karras hierarchy) and K5 (refit): every maximal subtree of the binary LBVH with
// at most `max_leaves` (<= 32) leaves is rebuilt top-down with the exact (full sweep) surface-area heuristic, one warp per
// subtree. Stands in for what the reference asks of the driver with ePreferFastTrace (reference main.cpp:419): the
// Morton order decides WHICH triangles share a subtree, the SAH decides HOW they are grouped inside it.
//
// Why this is cheap to splice into the LBVH: the internal nodes of the Karras subtree over the sorted leaves [a, b] are
// exactly the ids [a+1, b] when its root is a left child (root id = b) and [a, b-1] when it is a right child or the
// tree's root (root id = a) — by induction over left[i] = gamma, right[i] = gamma + 1 (k_lbvh_hierarchy) — so the rebuilt
// topology takes its node ids from that range, the root keeps its id, and nothing outside the subtree changes. The
// leaves are permuted inside [a, b] only: first[]/last[] of every node stay contiguous ranges, which is what the
// collapse (K6) needs. The permutation lives in leaf_prim[] (leaf position -> primitive); the sorted keys stay as
// they are for bpt_download_morton.
//
// One warp per subtree, lane k holds the leaf at position a + k (box, primitive id). A split of the range [s, e):
// for each axis, rank the centroids (ties by position), bring the boxes into rank order with shuffles, segmented
// prefix / suffix box scans, cost(p) = A(left) * |left| + A(right) * |right| for every split position, warp-minimum
// over positions and axes; then the items move to the order of the winning axis. An explicit stack of ranges in shared
// memory replaces the recursion.
#include "build.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarpsPerBlock = 8;

struct Box { float lx, ly, lz, hx, hy, hz; };
__device__ __forceinline__ Box box_shfl(const Box& b, int src) {
    return {__shfl_sync(FULL, b.lx, src), __shfl_sync(FULL, b.ly, src), __shfl_sync(FULL, b.lz, src),
            __shfl_sync(FULL, b.hx, src), __shfl_sync(FULL, b.hy, src), __shfl_sync(FULL, b.hz, src)};
}
__device__ __forceinline__ Box box_merge(const Box& a, const Box& b) {
    return {fminf(a.lx, b.lx), fminf(a.ly, b.ly), fminf(a.lz, b.lz), fmaxf(a.hx, b.hx), fmaxf(a.hy, b.hy), fmaxf(a.hz, b.hz)};
}
__device__ __forceinline__ float box_half_area(const Box& b) {
    const float dx = fmaxf(b.hx - b.lx, 0.f), dy = fmaxf(b.hy - b.ly, 0.f), dz = fmaxf(b.hz - b.lz, 0.f);  // empty box: no area
    return dx * dy + dy * dz + dz * dx;
}

// roots of the maximal subtrees with 3..max_leaves leaves (2 leaves have one topology)
__global__ void k_sah_roots(uint32_t n, uint32_t max_leaves, const uint32_t* __restrict__ parent,
                            const uint32_t* __restrict__ first, const uint32_t* __restrict__ last,
                            uint32_t* __restrict__ roots, uint32_t* __restrict__ nroots) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint32_t cnt = last[i] - first[i] + 1u;
    if (cnt < 3u || cnt > max_leaves) return;
    const uint32_t p = parent[i];
    if (p != 0xffffffffu && last[p] - first[p] + 1u <= max_leaves) return;  // the parent's subtree is rebuilt as a whole
    roots[atomicAdd(nroots, 1u)] = i;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) k_sah_rebuild(uint32_t n, const uint32_t* __restrict__ roots,
                                                                      const uint32_t* __restrict__ nroots,
                                                                      const float4* __restrict__ plo, const float4* __restrict__ phi,
                                                                      uint32_t* __restrict__ leaf_prim, uint32_t* __restrict__ left,
                                                                      uint32_t* __restrict__ right, uint32_t* __restrict__ parent,
                                                                      uint32_t* __restrict__ first, uint32_t* __restrict__ last) {
    __shared__ int s_inv[kWarpsPerBlock][32];       // rank -> lane of the current split
    __shared__ uint32_t s_stack[kWarpsPerBlock][32][2];  // {start | end << 8, node id}
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * kWarpsPerBlock + wib;
    if (w >= *nroots) return;
    const uint32_t root = roots[w];
    const uint32_t a = first[root], b = last[root];
    const int m = (int)(b - a + 1u);
    // the subtree's node ids are [a+1, b] under a left-child root (id b) and [a, b-1] otherwise (id a): the root keeps
    // its id and the m - 2 other inner nodes take [a+1, b-1] in the order they are created
    uint32_t next_id = a + 1u;

    uint32_t prim = 0;
    Box bx{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (lane < m) {
        prim = leaf_prim[a + lane];
        const float4 lo = plo[prim], hi = phi[prim];
        bx = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
    }
    int sp = 0;
    if (lane == 0) { s_stack[wib][0][0] = 0u | ((uint32_t)m << 8); s_stack[wib][0][1] = root; }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        --sp;
        const uint32_t se = s_stack[wib][sp][0], id = s_stack[wib][sp][1];
        __syncwarp();
        const int s = (int)(se & 0xffu), e = (int)(se >> 8);
        const int cnt = e - s;
        const bool in = lane >= s && lane < e;
        int best_pos = s;  // split after sorted position best_pos (left = [s, best_pos])
        int rank_keep = lane - s;
        if (cnt > 2) {
            float best_cost = 3.0e38f;
            const float cen[3] = {bx.lx + bx.hx, bx.ly + bx.hy, bx.lz + bx.hz};
#pragma unroll
            for (int axis = 0; axis < 3; ++axis) {
                const float key = cen[axis] == cen[axis] ? cen[axis] : 0.0f;  // a NaN would break the total order of the ranks
                int rank = 0;
                for (int j = s; j < e; ++j) {
                    const float kj = __shfl_sync(FULL, key, j);
                    rank += (kj < key || (kj == key && j < lane)) ? 1 : 0;
                }
                if (in) s_inv[wib][s + rank] = lane;
                __syncwarp();
                const int src = in ? s_inv[wib][lane] : lane;
                __syncwarp();
                const Box sb = box_shfl(bx, src);  // the box with rank lane - s
                Box pre = sb, suf = sb;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const Box up = box_shfl(pre, max(lane - d, 0));
                    const Box dn = box_shfl(suf, min(lane + d, 31));
                    if (in && lane - d >= s) pre = box_merge(pre, up);
                    if (in && lane + d < e) suf = box_merge(suf, dn);
                }
                const Box suf1 = box_shfl(suf, min(lane + 1, 31));  // suffix box starting one position to the right
                float cost = 3.0e38f;
                if (in && lane < e - 1)
                    cost = box_half_area(pre) * (float)(lane - s + 1) + box_half_area(suf1) * (float)(e - 1 - lane);
                // warp minimum; on ties (coincident primitives) the split closest to the middle, then the lowest position
                float c = cost;
                int pos = lane;
                const int mid2 = s + e - 2;  // twice the middle split position
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    const float oc = __shfl_xor_sync(FULL, c, d);
                    const int op = __shfl_xor_sync(FULL, pos, d);
                    const int ob = abs(2 * op - mid2), mb = abs(2 * pos - mid2);
                    if (oc < c || (oc == c && (ob < mb || (ob == mb && op < pos)))) { c = oc; pos = op; }
                }
                if (c < best_cost) { best_cost = c; best_pos = pos; rank_keep = rank; }
            }
            // move the items of [s, e) into the order of the winning axis
            if (in) s_inv[wib][s + rank_keep] = lane;
            __syncwarp();
            const int src = in ? s_inv[wib][lane] : lane;
            __syncwarp();
            bx = box_shfl(bx, src);
            prim = __shfl_sync(FULL, prim, src);
        }
        const int mid = cnt == 2 ? s + 1 : best_pos + 1;  // left = [s, mid), right = [mid, e)
        // children: a single leaf is node n - 1 + position; an inner node takes the next free id of the subtree's range
        const bool inner_l = mid - s > 1, inner_r = e - mid > 1;
        const uint32_t lc = inner_l ? next_id : n - 1u + a + (uint32_t)s;
        next_id += inner_l ? 1u : 0u;
        const uint32_t rc = inner_r ? next_id : n - 1u + a + (uint32_t)mid;
        next_id += inner_r ? 1u : 0u;
        if (lane == 0) {
            left[id] = lc; right[id] = rc;
            first[id] = a + (uint32_t)s; last[id] = a + (uint32_t)e - 1u;
            parent[lc] = id; parent[rc] = id;
            int top = sp;
            if (inner_l) { s_stack[wib][top][0] = (uint32_t)s | ((uint32_t)mid << 8); s_stack[wib][top][1] = lc; ++top; }
            if (inner_r) { s_stack[wib][top][0] = (uint32_t)mid | ((uint32_t)e << 8); s_stack[wib][top][1] = rc; }
        }
        sp += (inner_l ? 1 : 0) + (inner_r ? 1 : 0);
        __syncwarp();
    }
    if (lane < m) leaf_prim[a + lane] = prim;
}

__global__ void k_leaf_order_from_keys(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ leaf_prim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) leaf_prim[i] = (uint32_t)(keys[i] & 0xffffffffu);
}

}  // namespace

void sah_launch_leaf_order(const uint64_t* keys, uint32_t n, uint32_t* leaf_prim, cudaStream_t st) {
    k_leaf_order_from_keys<<<(n + 255u) / 256u, 256, 0, st>>>(keys, n, leaf_prim);
}

// roots / nroots: scratch of n + 1 uint32 (nroots zeroed here). max_leaves in [3, 32].
void sah_launch_rebuild(uint32_t n, uint32_t max_leaves, const float4* plo, const float4* phi, uint32_t* leaf_prim,
                        uint32_t* left, uint32_t* right, uint32_t* parent, uint32_t* first, uint32_t* last, uint32_t* roots,
                        uint32_t* nroots, cudaStream_t st) {
    if (n < 3) return;
    cudaMemsetAsync(nroots, 0, sizeof(uint32_t), st);
    k_sah_roots<<<(n + 255u) / 256u, 256, 0, st>>>(n, max_leaves, parent, first, last, roots, nroots);
    // at most (n - 1) / 2 roots (every root has >= 3 leaves and the subtrees are disjoint): n / 3 warps suffice
    const uint32_t max_roots = n / 3u + 1u;
    k_sah_rebuild<<<(max_roots + kWarpsPerBlock - 1) / kWarpsPerBlock, 32 * kWarpsPerBlock, 0, st>>>(
        n, roots, nroots, plo, phi, leaf_prim, left, right, parent, first, last);
}
