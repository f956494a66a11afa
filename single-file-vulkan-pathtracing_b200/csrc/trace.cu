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
//
// Instruction budget. The kernel is issue-bound (profiles/r1a_*: 72 % issue-active, ~0 DRAM
// stall), and on sm_100 the ALU pipe (PRMT/LOP3/FMNMX/SEL) runs at half the rate of the FMA pipe,
// so the node step is written to spend as few ALU-pipe instructions as possible:
//   * a quantised plane byte q becomes the float 32768+q with ONE byte-permute (magic word from the
//     constant bank, selector immediate); the 32768 bias is folded into the per-axis offset with one
//     FMA per axis instead of one subtraction per plane (the fold costs <= 2^-9 of a quantisation
//     step; the builder quantises with a 2^-7-step margin, build.cu)
//   * a child that passes the slab test ORs one immediate into the hit word (its internal-child bit
//     and its three triangle bits); one AND with the node's `valid` word then yields the internal
//     hits (top byte, slot order) and the triangle hits (low 24 bits) — no per-child meta decoding
//   * the octant permutation of the internal hits (bit p <- slot p ^ oct) is one byte lookup in a
//     2 KB shared-memory table (LSU pipe) instead of three conditional bit-swap stages
//   * triangles are software-pipelined: every loop iteration does at most ONE node step and ONE
//     triangle test per lane; a lane keeps descending (and popping node groups) while its triangle
//     group drains, so the triangle test runs once per iteration with every lane that has a
//     pending triangle instead of a divergent inner loop that ran at 2/32 lanes
#include "trace.cuh"

namespace {

constexpr int kRefillBelow = 24;   // refill when fewer live lanes than this (>= 9 rays per atomic)
constexpr int kLocalStack = kTraceLocalStack;
constexpr int kStepsPerRefill = 4; // traversal iterations between two refill votes
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

// Byte J of w as the float 32768 + byte (exact): bytes {0, w.bJ, 0, 0x47}. `magic` (0x47000000) arrives as a kernel
// argument so that it stays a constant-bank operand and the selector is the instruction's immediate — when both are
// compile-time constants ptxas keeps the magic as the immediate and burns a register move per selector.
template <int J>
__device__ __forceinline__ float byte_f(uint32_t w, uint32_t magic) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7604 + 0x10 * J));
    return __uint_as_float(r);
}
constexpr float kByteBias = 32768.0f;

struct RayState {
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, tmin, tbest;
    uint32_t htri;    // leaf slot of the closest hit, BPT_MISS if none
    uint32_t oct;     // dx>=0?4:0 | dy>=0?2:0 | dz>=0?1:0: slot s is visited with priority s ^ oct
};

// Tests the 4 children (slots 4Q..4Q+3) of one half of a node; returns the hit word contributions: bit 24+s and
// bits 3s..3s+2 for every slot s whose box the ray segment overlaps. bx/by/bz carry the -32768*ad bias of byte_f.
template <int Q>
__device__ __forceinline__ uint32_t test_quad(uint32_t xn, uint32_t yn, uint32_t zn, uint32_t xf, uint32_t yf,
                                              uint32_t zf, float adx, float ady, float adz, float bx, float by,
                                              float bz, float tmin, float tbest, uint32_t magic) {
    uint32_t hit = 0;
#define BPT_CHILD(J)                                                                                        \
    {                                                                                                       \
        float tx0 = fmaf(byte_f<J>(xn, magic), adx, bx), tx1 = fmaf(byte_f<J>(xf, magic), adx, bx);         \
        float ty0 = fmaf(byte_f<J>(yn, magic), ady, by), ty1 = fmaf(byte_f<J>(yf, magic), ady, by);         \
        float tz0 = fmaf(byte_f<J>(zn, magic), adz, bz), tz1 = fmaf(byte_f<J>(zf, magic), adz, bz);         \
        float tn = fmaxf(fmaxf(tx0, ty0), fmaxf(tz0, tmin));                                                \
        float tf = fminf(fminf(tx1, ty1), fminf(tz1, tbest));                                               \
        if (tn <= tf) hit |= (7u << (3 * (4 * Q + J))) | (1u << (24 + 4 * Q + J));                          \
    }
    BPT_CHILD(0) BPT_CHILD(1) BPT_CHILD(2) BPT_CHILD(3)
#undef BPT_CHILD
    return hit;
}

template <int BLOCK, int SSTACK, bool COUNT>
__global__ void __launch_bounds__(BLOCK, 1) k_trace(TraceArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint8_t* slut = smem_raw + 16;  // slut[oct * 256 + b]: bit p = bit (p ^ oct) of b
    uint2* sstack = reinterpret_cast<uint2*>(smem_raw + kTraceSmemFixed);
    uint4* snodes = reinterpret_cast<uint4*>(smem_raw + kTraceSmemFixed + (size_t)SSTACK * BLOCK * sizeof(uint2));
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

    for (uint32_t i = threadIdx.x; i < 2048u; i += BLOCK) {
        const uint32_t oct = i >> 8, b = i & 0xffu;
        uint32_t p = 0;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) p |= ((b >> (k ^ oct)) & 1u) << k;
        slut[i] = (uint8_t)p;
    }
    __syncthreads();

    if (blockIdx.x == 0 && threadIdx.x == 0 && a.stat) atomicAdd(a.stat + BPT_STAT_RAYS, (unsigned long long)nrays);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    uint2* mystack = sstack + threadIdx.x;
    uint2 lstack[kLocalStack];
    const uint32_t magic = a.magic;

    RayState r;
    uint2 G = make_uint2(0u, 0u);  // node group: x = first internal child, y = hits by priority << 24 | internal mask
    uint2 T = make_uint2(0u, 0u);  // triangle group: x = node the triangles belong to, y = hit bits (valid layout)
    uint32_t Tb = 0u, Tv = 0u;     // tri_base and valid word of node T.x
    int sp = 0;
    uint32_t ray_idx = 0;
    bool active = false, exhausted = false;
    unsigned long long cnt_nodes = 0, cnt_tris = 0, cnt_witer = 0, cnt_wnode = 0, cnt_wtri = 0, cnt_liter = 0;

#define BPT_PUSH(E)                                                               \
    {                                                                             \
        if (sp < SSTACK) mystack[sp * BLOCK] = (E); else lstack[sp - SSTACK] = (E); \
        ++sp;                                                                     \
    }
#define BPT_LEADER() ((__activemask() & lt) == 0u)

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
                    r.oct = (rd.x >= 0.f ? 4u : 0u) | (rd.y >= 0.f ? 2u : 0u) | (rd.z >= 0.f ? 1u : 0u);
                    r.htri = BPT_MISS;
                    G = make_uint2(0u, 0x80000000u);  // root: node 0 through priority bit 31, internal mask 0
                    T = make_uint2(0u, 0u);
                    sp = 0;
                    active = true;
                }
            }
            if (base + (uint32_t)nidle >= nrays) exhausted = true;
            actmask = __ballot_sync(FULL, active);
        }
        if (actmask == 0u) break;

        // Every lane of the warp runs the loop below, live or not (an idle lane has no node bits, no triangle bits and
        // an empty stack), and the warp reconverges with __syncwarp() in front of every phase: without it the lanes
        // that popped and the lanes that did not reach the node step as two separate groups, and the ~270-instruction
        // node step runs twice per iteration at 12/32 lanes (measured; profiles/).
        {
#pragma unroll 1
            for (int it = 0; it < kStepsPerRefill; ++it) {
                if (COUNT) {
                    const unsigned m = __ballot_sync(FULL, active);
                    if (lane == 0 && m) ++cnt_witer;
                    if (active) ++cnt_liter;
                }
                // ---------------- pop when out of node work: a node group always, a parked triangle group only when
                //                  none is draining (it stays on the stack until then)
                if (active && !(G.y & 0xff000000u) && sp > 0) {
                    const uint2 e = sp <= SSTACK ? mystack[(sp - 1) * BLOCK] : lstack[sp - 1 - SSTACK];
                    if (e.y & 0xff000000u) { G = e; --sp; }
                    else if (T.y == 0u) {
                        T = e; --sp;
                        const uint4 w1 = e.x < a.top_nodes ? snodes[5 * (size_t)e.x + 1] : __ldg(a.nodes + 5 * (size_t)e.x + 1);
                        Tb = w1.y; Tv = w1.z;
                    }
                }
                // ---------------- node step
                __syncwarp();
                if (G.y & 0xff000000u) {
                    if (COUNT) { if (BPT_LEADER()) ++cnt_wnode; ++cnt_nodes; }
                    const uint32_t bit = 31u - __clz(G.y);
                    G.y &= ~(1u << bit);
                    if (G.y & 0xff000000u) BPT_PUSH(G)  // siblings still pending
                    const uint32_t slot = (bit - 24u) ^ r.oct;
                    const uint32_t rel = __popc(G.y & ~(0xffffffffu << slot) & 0xffu);
                    const uint32_t node = G.x + rel;
                    uint4 n0, n1, n2, n3, n4;
                    if (node < a.top_nodes) {  // staged prefix: shared memory
                        const uint4* np = snodes + 5 * (size_t)node;
                        n0 = np[0]; n1 = np[1]; n2 = np[2]; n3 = np[3]; n4 = np[4];
                    } else {
                        const uint4* np = a.nodes + 5 * (size_t)node;
                        n0 = __ldg(np); n1 = __ldg(np + 1); n2 = __ldg(np + 2); n3 = __ldg(np + 3); n4 = __ldg(np + 4);
                    }
                    const float adx = __uint_as_float((n0.w & 0xffu) << 23) * r.idx;
                    const float ady = __uint_as_float(((n0.w >> 8) & 0xffu) << 23) * r.idy;
                    const float adz = __uint_as_float(((n0.w >> 16) & 0xffu) << 23) * r.idz;
                    const float bx = fmaf(adx, -kByteBias, (__uint_as_float(n0.x) - r.ox) * r.idx);
                    const float by = fmaf(ady, -kByteBias, (__uint_as_float(n0.y) - r.oy) * r.idy);
                    const float bz = fmaf(adz, -kByteBias, (__uint_as_float(n0.z) - r.oz) * r.idz);
                    // near / far plane bytes by direction sign
                    const bool nx = r.idx < 0.f, ny = r.idy < 0.f, nz = r.idz < 0.f;
                    const uint32_t xn0 = nx ? n3.z : n2.x, xn1 = nx ? n3.w : n2.y, xf0 = nx ? n2.x : n3.z, xf1 = nx ? n2.y : n3.w;
                    const uint32_t yn0 = ny ? n4.x : n2.z, yn1 = ny ? n4.y : n2.w, yf0 = ny ? n2.z : n4.x, yf1 = ny ? n2.w : n4.y;
                    const uint32_t zn0 = nz ? n4.z : n3.x, zn1 = nz ? n4.w : n3.y, zf0 = nz ? n3.x : n4.z, zf1 = nz ? n3.y : n4.w;
                    uint32_t hit = test_quad<0>(xn0, yn0, zn0, xf0, yf0, zf0, adx, ady, adz, bx, by, bz, r.tmin, r.tbest, magic);
                    hit |= test_quad<1>(xn1, yn1, zn1, xf1, yf1, zf1, adx, ady, adz, bx, by, bz, r.tmin, r.tbest, magic);
                    hit &= n1.z;  // valid: internal children in the top byte, triangles in the low 24 bits
                    // internal hits in visiting priority (bit p <- slot p ^ oct), internal mask in the low byte
                    const uint32_t prio = slut[(r.oct << 8) + (hit >> 24)];
                    G.x = n1.x;
                    G.y = __byte_perm(prio, n1.z, 0x0217);
                    if (hit & 0x00ffffffu) {
                        if (T.y) BPT_PUSH(T)  // a group is still draining: park it, it is popped like any other entry
                        T.x = node;
                        T.y = hit & 0x00ffffffu;
                        Tb = n1.y; Tv = n1.z;
                    }
                }
                // ---------------- triangle step: one triangle of the lane's pending group
                __syncwarp();
                if (T.y) {
                    if (COUNT) { if (BPT_LEADER()) ++cnt_wtri; ++cnt_tris; }
                    const uint32_t k = 31u - __clz(T.y);
                    T.y &= ~(1u << k);
                    const uint32_t tri = Tb + __popc(Tv & ~(0xffffffffu << k));  // k < 24: internal bits never counted
                    float4 ru, rv, rw;
                    if (tri < a.top_tris) {
                        const float4* tp = stris + 3 * (size_t)tri;
                        ru = tp[0]; rv = tp[1]; rw = tp[2];
                    } else {
                        const float4* tp = a.woop + 3 * (size_t)tri;
                        ru = __ldg(tp); rv = __ldg(tp + 1); rw = __ldg(tp + 2);
                    }
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
                            if (take) { r.tbest = t; r.htri = tri; }
                        }
                    }
                }
                // ---------------- terminate
                __syncwarp();
                if (active && !(G.y & 0xff000000u) && T.y == 0u && sp == 0) {
                    uint4 h;  // u, v are re-derived from the original vertices by the consumer (shade.cu barycentrics)
                    h.x = __float_as_uint(r.tbest);
                    h.y = 0u;
                    h.z = 0u;
                    h.w = r.htri == BPT_MISS ? BPT_MISS : __ldg(&a.prim_index[r.htri]);
                    a.hits[ray_idx] = h;
                    active = false;
                }
            }
        }
    }
#undef BPT_PUSH
#undef BPT_LEADER
    if (COUNT) {
        unsigned long long c[6] = {cnt_nodes, cnt_tris, cnt_witer, cnt_wnode, cnt_wtri, cnt_liter};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int o = 16; o; o >>= 1) c[i] += __shfl_xor_sync(FULL, c[i], o);
            if (lane == 0) atomicAdd(a.stat + BPT_STAT_NODES + i, c[i]);
        }
    }
}

}  // namespace

// shared memory the traversal kernel needs for a given staging configuration
size_t trace_smem_bytes(uint32_t top_nodes, uint32_t top_tris) {
    return kTraceSmemFixed + (size_t)kTraceSmemStack * kTraceBlock * sizeof(uint2) + (size_t)top_nodes * 80 + (size_t)top_tris * 48;
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
