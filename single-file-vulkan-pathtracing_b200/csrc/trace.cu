// trace.cu — K10: persistent-threads closest-hit traversal of the compressed BVH8.
//
// Replaces traceRayEXT (reference shaders/raygen.rgen:63-75: closest opaque hit in [tmin,tmax],
// cull mask 0xff, no face culling per main.cpp:525) and the fixed-function traversal +
// ray/triangle test of the driver's acceleration structure.
//
// Design (B200):
//   * one persistent CTA per SM (grid = #SMs * ctas_per_sm); warps pull rays from a global counter
//     and refill idle lanes (ballot + popc prefix) whenever fewer than kRefillBelow lanes are live
//   * the BFS prefix of the node array (top of the tree) and the first triangles are staged once
//     per CTA into shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//     deeper nodes / triangles are read with 128-bit loads through L1/L2
//   * per-lane traversal stack: kSmemStack 8-byte entries in shared memory (thread-interleaved,
//     conflict-free), overflow in local memory
//   * rays are 2 x float4, hits 1 x uint4 -> all record traffic is 128-bit
//   * traversal order: octant-permuted slot priority (Ylitie et al. 2017), node groups and
//     triangle groups share one 64-bit stack entry format
#include "trace.cuh"

namespace {

constexpr int kRefillBelow = 24;   // refill when fewer live lanes than this (>= 9 rays per atomic)
constexpr int kLocalStack = 40;    // overflow entries in local memory
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------- TMA / mbarrier (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 0xff in every byte of x whose top bit is set. __byte_perm() masks each selector nibble to 3 bits, which drops
// prmt's sign-replicate mode (selector bit 3), so the instruction is issued directly.
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
}

// byte j of w as float (exact): build 2^23 + byte with a byte permute, subtract 2^23
template <int J>
__device__ __forceinline__ float byte_f(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + J)) - 8388608.0f;
}

struct RayState {
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, tmin, tbest, hu, hv;
    uint32_t htri;    // leaf slot of the closest hit, BPT_MISS if none
    uint32_t octinv4; // (dx>=0?4:0 | dy>=0?2:0 | dz>=0?1:0) replicated in 4 bytes
};

// Tests the 4 children of one half of a node; returns their contribution to the hit mask.
__device__ __forceinline__ uint32_t test_quad(uint32_t meta4, uint32_t xn, uint32_t yn, uint32_t zn, uint32_t xf,
                                              uint32_t yf, uint32_t zf, float adx, float ady, float adz, float bx,
                                              float by, float bz, float tmin, float tbest, uint32_t octinv4) {
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);  // 0xff where internal
    const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    uint32_t hit = 0;
#define BPT_CHILD(J)                                                                            \
    {                                                                                           \
        float tx0 = fmaf(byte_f<J>(xn), adx, bx), tx1 = fmaf(byte_f<J>(xf), adx, bx);           \
        float ty0 = fmaf(byte_f<J>(yn), ady, by), ty1 = fmaf(byte_f<J>(yf), ady, by);           \
        float tz0 = fmaf(byte_f<J>(zn), adz, bz), tz1 = fmaf(byte_f<J>(zf), adz, bz);           \
        float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin));                                    \
        float tf = fminf(fminf(tx1, ty1), fminf(tz1, tbest));                                   \
        if (tn <= tf) hit |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu); \
    }
    BPT_CHILD(0) BPT_CHILD(1) BPT_CHILD(2) BPT_CHILD(3)
#undef BPT_CHILD
    return hit;
}

template <int BLOCK, int SSTACK, bool COUNT>
__global__ void __launch_bounds__(BLOCK, 1) k_trace(TraceArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint2* sstack = reinterpret_cast<uint2*>(smem_raw + 16);
    uint4* snodes = reinterpret_cast<uint4*>(smem_raw + 16 + (size_t)SSTACK * BLOCK * sizeof(uint2));
    float4* stris = reinterpret_cast<float4*>(snodes + 5 * (size_t)a.top_nodes);

    const uint32_t nrays = *a.count_ptr;
    if (nrays == 0u) return;  // uniform across the grid: an exhausted bounce costs one launch and nothing else

    // ---- stage the top of the tree into shared memory with TMA bulk copies
    const uint32_t node_bytes = a.top_nodes * 80u, tri_bytes = a.top_tris * 48u;
    if (node_bytes + tri_bytes) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, node_bytes + tri_bytes);
            constexpr uint32_t kChunk = 32768u;
            for (uint32_t off = 0; off < node_bytes; off += kChunk)
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(snodes) + off,
                             reinterpret_cast<const unsigned char*>(a.nodes) + off, min(kChunk, node_bytes - off), bar);
            for (uint32_t off = 0; off < tri_bytes; off += kChunk)
                tma_bulk_g2s(reinterpret_cast<unsigned char*>(stris) + off,
                             reinterpret_cast<const unsigned char*>(a.woop) + off, min(kChunk, tri_bytes - off), bar);
        }
        mbar_wait(bar, 0);
    }

    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stat_rays) atomicAdd(a.stat_rays, (unsigned long long)nrays);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    uint2* mystack = sstack + threadIdx.x;
    uint2 lstack[kLocalStack];

    RayState r;
    uint2 G = make_uint2(0u, 0u), T = make_uint2(0u, 0u);
    int sp = 0;
    uint32_t ray_idx = 0;
    bool active = false, exhausted = false;
    unsigned long long cnt_nodes = 0, cnt_tris = 0;

    for (;;) {
        unsigned actmask = __ballot_sync(FULL, active);
        if (!exhausted && __popc(actmask) < kRefillBelow) {
            const unsigned idle = ~actmask;
            const int nidle = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.fetch_ctr, (uint32_t)nidle);
            base = __shfl_sync(FULL, base, 0);
            if (!active) {
                uint32_t ri = base + __popc(idle & lt);
                if (ri < nrays) {
                    const float4 ro = __ldg(&a.rays[2 * (size_t)ri]);
                    const float4 rd = __ldg(&a.rays[2 * (size_t)ri + 1]);
                    ray_idx = ri;
                    r.ox = ro.x; r.oy = ro.y; r.oz = ro.z; r.tmin = ro.w;
                    r.dx = rd.x; r.dy = rd.y; r.dz = rd.z; r.tbest = rd.w;
                    const float eps = 1e-30f;  // keep 1/d finite; direction sign is kept
                    r.idx = 1.0f / (fabsf(rd.x) > eps ? rd.x : copysignf(eps, rd.x));
                    r.idy = 1.0f / (fabsf(rd.y) > eps ? rd.y : copysignf(eps, rd.y));
                    r.idz = 1.0f / (fabsf(rd.z) > eps ? rd.z : copysignf(eps, rd.z));
                    r.octinv4 = ((rd.x >= 0.f ? 4u : 0u) | (rd.y >= 0.f ? 2u : 0u) | (rd.z >= 0.f ? 1u : 0u)) * 0x01010101u;
                    r.htri = BPT_MISS; r.hu = 0.f; r.hv = 0.f;
                    G = make_uint2(0u, 0x80000000u);  // root: node 0 through priority bit 31, imask 0
                    T = make_uint2(0u, 0u);
                    sp = 0;
                    active = true;
                }
            }
            if (base + (uint32_t)nidle >= nrays) exhausted = true;
            actmask = __ballot_sync(FULL, active);
        }
        if (actmask == 0u) break;

        if (active) {
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                // ---------------- node step
                if (G.y & 0xff000000u) {
                    const uint32_t bit = 31u - __clz(G.y);
                    G.y &= ~(1u << bit);
                    if (G.y & 0xff000000u) {  // siblings still pending: push them
                        if (sp < SSTACK) mystack[sp * BLOCK] = G; else lstack[sp - SSTACK] = G;
                        ++sp;
                    }
                    const uint32_t slot = (bit - 24u) ^ (r.octinv4 & 7u);
                    const uint32_t rel = __popc(G.y & ~(0xffffffffu << slot) & 0xffu);
                    const uint32_t node = G.x + rel;
                    // generic pointer: shared window for the staged prefix, global otherwise
                    const uint4* np = node < a.top_nodes ? snodes + 5 * (size_t)node : a.nodes + 5 * (size_t)node;
                    const uint4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3], n4 = np[4];
                    if (COUNT) ++cnt_nodes;
                    const float adx = __uint_as_float((n0.w & 0xffu) << 23) * r.idx;
                    const float ady = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * r.idy;
                    const float adz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * r.idz;
                    const float bx = (__uint_as_float(n0.x) - r.ox) * r.idx;
                    const float by = (__uint_as_float(n0.y) - r.oy) * r.idy;
                    const float bz = (__uint_as_float(n0.z) - r.oz) * r.idz;
                    // near / far plane bytes by direction sign
                    const bool nx = r.idx < 0.f, ny = r.idy < 0.f, nz = r.idz < 0.f;
                    const uint32_t xn0 = nx ? n3.z : n2.x, xn1 = nx ? n3.w : n2.y, xf0 = nx ? n2.x : n3.z, xf1 = nx ? n2.y : n3.w;
                    const uint32_t yn0 = ny ? n4.x : n2.z, yn1 = ny ? n4.y : n2.w, yf0 = ny ? n2.z : n4.x, yf1 = ny ? n2.w : n4.y;
                    const uint32_t zn0 = nz ? n4.z : n3.x, zn1 = nz ? n4.w : n3.y, zf0 = nz ? n3.x : n4.z, zf1 = nz ? n3.y : n4.w;
                    uint32_t hitmask = test_quad(n1.z, xn0, yn0, zn0, xf0, yf0, zf0, adx, ady, adz, bx, by, bz, r.tmin, r.tbest, r.octinv4);
                    hitmask |= test_quad(n1.w, xn1, yn1, zn1, xf1, yf1, zf1, adx, ady, adz, bx, by, bz, r.tmin, r.tbest, r.octinv4);
                    G.x = n1.x;
                    G.y = (hitmask & 0xff000000u) | (n0.w >> 24);
                    T.x = n1.y;
                    T.y = hitmask & 0x00ffffffu;
                } else {
                    T = G;
                    G = make_uint2(0u, 0u);
                }
                // ---------------- triangle steps
                while (T.y) {
                    const uint32_t k = 31u - __clz(T.y);
                    T.y &= ~(1u << k);
                    const uint32_t tri = T.x + k;
                    const float4* tp = tri < a.top_tris ? stris + 3 * (size_t)tri : a.woop + 3 * (size_t)tri;
                    const float4 ru = tp[0], rv = tp[1], rw = tp[2];
                    if (COUNT) ++cnt_tris;
                    const float oz = rw.w + r.ox * rw.x + r.oy * rw.y + r.oz * rw.z;
                    const float dz = r.dx * rw.x + r.dy * rw.y + r.dz * rw.z;
                    const float t = __fdividef(-oz, dz);
                    if (t >= r.tmin && t <= r.tbest) {
                        const float ou = ru.w + r.ox * ru.x + r.oy * ru.y + r.oz * ru.z;
                        const float du = r.dx * ru.x + r.dy * ru.y + r.dz * ru.z;
                        const float u = ou + t * du;
                        const float ov = rv.w + r.ox * rv.x + r.oy * rv.y + r.oz * rv.z;
                        const float dv = r.dx * rv.x + r.dy * rv.y + r.dz * rv.z;
                        const float v = ov + t * dv;
                        if (u >= 0.f && v >= 0.f && u + v <= 1.f) {
                            // equal distance (exact duplicate triangles): lowest primitive id wins
                            bool take = t < r.tbest || r.htri == BPT_MISS;
                            if (!take) take = __ldg(&a.prim_index[tri]) < __ldg(&a.prim_index[r.htri]);
                            if (take) { r.tbest = t; r.hu = u; r.hv = v; r.htri = tri; }
                        }
                    }
                }
                // ---------------- pop / terminate
                if (!(G.y & 0xff000000u)) {
                    if (sp == 0) {
                        uint4 h;
                        h.x = __float_as_uint(r.tbest);
                        h.y = __float_as_uint(r.hu);
                        h.z = __float_as_uint(r.hv);
                        h.w = r.htri == BPT_MISS ? BPT_MISS : __ldg(&a.prim_index[r.htri]);
                        a.hits[ray_idx] = h;
                        active = false;
                        break;
                    }
                    --sp;
                    G = sp < SSTACK ? mystack[sp * BLOCK] : lstack[sp - SSTACK];
                }
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            cnt_nodes += __shfl_xor_sync(FULL, cnt_nodes, o);
            cnt_tris += __shfl_xor_sync(FULL, cnt_tris, o);
        }
        if (lane == 0) {
            atomicAdd(a.stat_nodes, cnt_nodes);
            atomicAdd(a.stat_tris, cnt_tris);
        }
    }
}

}  // namespace

// shared memory the traversal kernel needs for a given staging configuration
size_t trace_smem_bytes(uint32_t top_nodes, uint32_t top_tris) {
    return 16 + (size_t)kTraceSmemStack * kTraceBlock * sizeof(uint2) + (size_t)top_nodes * 80 + (size_t)top_tris * 48;
}

cudaError_t trace_configure() {
    cudaError_t e = cudaFuncSetAttribute(k_trace<kTraceBlock, kTraceSmemStack, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kTraceMaxSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_trace<kTraceBlock, kTraceSmemStack, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kTraceMaxSmem);
}

void trace_launch(const TraceArgs& a, unsigned grid, bool count, cudaStream_t st) {
    size_t smem = trace_smem_bytes(a.top_nodes, a.top_tris);
    if (count) k_trace<kTraceBlock, kTraceSmemStack, true><<<grid, kTraceBlock, smem, st>>>(a);
    else k_trace<kTraceBlock, kTraceSmemStack, false><<<grid, kTraceBlock, smem, st>>>(a);
}
