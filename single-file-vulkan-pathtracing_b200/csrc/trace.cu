// trace.cu — K10: persistent-threads closest-hit traversal of the compressed BVH8.
//
// Replaces traceRayEXT (reference shaders/raygen.rgen:63-75: closest opaque hit in [tmin,tmax],
// cull mask 0xff, no face culling per main.cpp:525) and the fixed-function traversal +
// ray/triangle test of the driver's acceleration structure.
//
// Design (B200):
//   * one persistent CTA of 32 warps per SM (64 registers/thread = the whole register file); every warp keeps a pool
//     of 32 rays in shared memory (one atomic on the global counter + two coalesced 128-bit loads per lane per pool;
//     1/d and the octant are computed there, at full SIMD width) and tops its idle lanes up from it (ballot + popc
//     prefix) whenever fewer than `refill_below` lanes are live
//   * per-lane traversal stack: kSmemStack 8-byte entries in shared memory (thread-interleaved, conflict-free),
//     overflow in local memory; node groups, parked triangle groups and the instance sentinel share the entry format
//   * rays are 2 x float4, hits 1 x uint4 -> all queue traffic is 128-bit
//   * traversal order: octant-permuted slot priority (Ylitie et al. 2017)
//
// What bounds it (profiles/). The kernel moves few DRAM bytes (3-12 % of peak from bounce 0 to 7, L2
// hit rate 89-61 %): instruction issue binds it (74-65 % of the slots busy, ALU pipe 72-62 %). With
// 80- and 96-byte nodes it was bound by the SM's L1 data pipe: every lane walks its own node, nothing coalesces, and
// every 32-byte sector a lane receives costs one cycle of that pipe (82-87 % busy; a 4th sector per node visit cost
// 23 %). With 64-byte records the bound is instruction issue (70-73 %) and the ALU pipe (PRMT/LOP3/FMNMX/SEL: half
// the rate of the FMA pipe on sm_100, 63-66 % busy). Hence:
//   * nodes and triangle records are 64 bytes = two 256-bit loads (LDG.E.256) in ONE record array: the node was
//     squeezed from Ylitie's 80 bytes to 64 (21-bit grid origin, one shared exponent, 2-bit triangle counts, one
//     child pointer for nodes and triangles alike; common.cuh)
//   * a quantised plane byte q becomes the float 32768+q with ONE byte-permute (magic word from the constant bank,
//     selector immediate); the 32768 bias is folded into the per-axis offset with one FMA per axis instead of one
//     subtraction per plane (the fold costs <= 2^-9 of a quantisation step; the builder quantises with a 2^-7-step
//     margin, build.cu); two planes share one packed FMA (FFMA2)
//   * a child that passes the slab test ORs one immediate into the hit word (its internal-child bit and a 2-bit
//     all-ones triangle field); one AND with the node's `valid` word then yields the internal hits (bits 16..23,
//     slot order) and, per hit leaf slot, its triangle count (bits 0..15) — no per-child meta decoding
//   * the octant permutation of the internal hits (bit p <- slot p ^ oct) is one 64-bit lookup in a 2 KB
//     shared-memory table (LSU pipe; the row of a hit byte holds its 8 permutations, so lanes with equal hit bytes
//     broadcast) instead of three conditional bit-swap stages
//   * triangles are software-pipelined: every loop iteration does at most ONE node step and ONE triangle test per
//     lane; a lane keeps descending (and popping node groups) while its triangle group drains, so the triangle test
//     runs once per iteration with every lane that has a pending triangle instead of a divergent inner loop that
//     ran at 2/32 lanes (the STAGED instance tests two per iteration: with the records in shared memory and a mesh of
//     36 triangles under 6 nodes a lane has more triangles than nodes to work through; +16 % on the Cornell box)
//   * the whole warp runs every phase of the loop behind a __syncwarp(): without it the lanes that popped and the
//     lanes that did not reach the node step as two groups and the node step runs twice per iteration at 12/32 lanes
//   * both halves of a triangle record are consumed in one basic block (distance and barycentrics evaluated together):
//     with the barycentrics behind the distance test ptxas sinks the first half's load into that branch and every
//     iteration pays a third exposed memory round trip (10 % of all stall samples; +5 % when removed)
//   * small scenes (Cornell box, or 1000 instances of it: every record fits) run the STAGED instance: the record
//     array is copied into shared memory once per CTA by TMA bulk copies (cp.async.bulk + mbarrier complete_tx) and
//     never touched in global memory again; big scenes run the global instance (staging only the top of the tree was
//     measured at +-1 %: the shared/global dual path costs the issue slots the saved round trips buy)
#include "shade_one.cuh"
#include "trace.cuh"

namespace {

constexpr int kLocalStack = kTraceLocalStack;
// (pop + node step) rounds per loop iteration: one triangle step, terminate test and loop overhead per BPT_NODE_STEPS
// node steps (compile-time: even a trip-count-1 loop changes the register allocation of the kernel)
// Experiment knob (compile time): the instance that reads its records from global memory runs BPT_TRACE_CTAS_G CTAs of
// BPT_TRACE_BLOCK_G threads per SM instead of one CTA of 1024 (e.g. 2 x 576 = 36 warps at 56 registers per thread)
#ifndef BPT_TRACE_BLOCK_G
#define BPT_TRACE_BLOCK_G kTraceBlock
#define BPT_TRACE_CTAS_G 1
#endif
#ifndef BPT_NODE_STEPS
#define BPT_NODE_STEPS 1
#endif
// 1: barycentrics only behind the distance test (the round-1 kernel; kept for the A/B in DESIGN.md)
#ifndef BPT_TRI_LAZY_UV
#define BPT_TRI_LAZY_UV 0
#endif
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

// Shared memory through 32-bit shared-window addresses: with C++ pointers into the dynamic array ptxas rebuilds the
// generic window base (S2R SR_CgaCtaId + LEA) in front of every access of the traversal loop.
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 as_float4(uint4 q) {
    return make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// 256-bit read-only global load (LDG.E.256, sm_100+): p is 32-byte aligned
struct U8 { uint4 lo, hi; };
__device__ __forceinline__ U8 ldg256(const void* p) {
    U8 v;
    asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"  // volatile: issue where written
        : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w)
        : "l"(p));
    return v;
}
__device__ __forceinline__ U8 lds256(uint32_t addr) { return U8{lds128(addr), lds128(addr + 16u)}; }

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

// 1/d for the slab tests: |d| is kept away from 0 (the sign survives) and the reciprocal is the hardware approximation
// (MUFU.RCP, ~1 ulp): its error is far inside the padding the builder gives every box (2^-20 of the scene scale).
__device__ __forceinline__ float safe_rcp(float d) {
    const float eps = 1e-30f;
    return __fdividef(1.0f, fabsf(d) > eps ? d : copysignf(eps, d));
}

// This file is compiled with --fmad=false (the fused instance shades paths with the shader's IEEE operation order,
// shade_one.cuh), so the traversal arithmetic names its fused multiply-adds itself.
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fmaf(az, bz, fmaf(ax, bx, ay * by));
}
__device__ __forceinline__ float dot3p(float ax, float ay, float az, float bx, float by, float bz, float c) {
    return fmaf(az, bz, fmaf(ay, by, fmaf(ax, bx, c)));
}

struct RayState {
    float ox, oy, oz, dx, dy, dz, idx, idy, idz, tmin, tbest;
    uint32_t hprim;   // primitive of the closest hit, BPT_MISS if none
    uint32_t oct;     // dx>=0?4:0 | dy>=0?2:0 | dz>=0?1:0: slot s is visited with priority s ^ oct
};

// Two slab planes at once: {a0, a1} * b + c with the packed FP32 FMA of sm_100 (FFMA2; b and c are scalar operands
// the instruction broadcasts). The kernel is issue-bound, and this halves the 48 FMAs of a node step.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
    uint64_t A, B, C, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(C) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(C));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}

// Tests the 4 children (slots 4Q..4Q+3) of one half of a node; returns the hit word contributions: bit 16+s and
// bits 2s..2s+1 for every slot s whose box the ray segment overlaps (the node's valid word keeps the internal-child
// bit of internal children and the triangle count of leaf children). bx/by/bz carry the -32768*ad bias of byte_f.
template <int Q>
__device__ __forceinline__ uint32_t test_quad(uint32_t xn, uint32_t yn, uint32_t zn, uint32_t xf, uint32_t yf,
                                              uint32_t zf, float adx, float ady, float adz, float bx, float by,
                                              float bz, float tmin, float tbest, uint32_t magic) {
    float tx0[4], tx1[4], ty0[4], ty1[4], tz0[4], tz1[4];
#define BPT_PLANES(T, W, AD, B)                                                        \
    fma2(T[0], T[1], byte_f<0>(W, magic), byte_f<1>(W, magic), AD, B);                 \
    fma2(T[2], T[3], byte_f<2>(W, magic), byte_f<3>(W, magic), AD, B);
    BPT_PLANES(tx0, xn, adx, bx) BPT_PLANES(tx1, xf, adx, bx)
    BPT_PLANES(ty0, yn, ady, by) BPT_PLANES(ty1, yf, ady, by)
    BPT_PLANES(tz0, zn, adz, bz) BPT_PLANES(tz1, zf, adz, bz)
#undef BPT_PLANES
    uint32_t hit = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float tn = fmaxf(fmaxf(tx0[j], ty0[j]), fmaxf(tz0[j], tmin));
        const float tf = fminf(fminf(tx1[j], ty1[j]), fminf(tz1[j], tbest));
        if (tn <= tf) hit |= (3u << (2 * (4 * Q + j))) | (1u << (16 + 4 * Q + j));
    }
    return hit;
}

// ---------------------------------------------------------------- fused path kernel: the shade batch
// Slot layout (64 bytes, shared memory): +0 {o.xyz, path id}  +16 {d.xyz, seed}  +32 {1/d.xyz, octant | depth << 8}
// +48 {throughput rgb, -}; a finished ray overwrites +32, +36 with {t, primitive}. A slot that has never held a path
// carries kNewPath as its primitive.
constexpr uint32_t kNewPath = 0xfffffffeu;
constexpr uint32_t kFusedReady = 64u, kFusedCtl = 128u, kFusedSlot0 = 192u;  // offsets inside a warp's block

// Shades up to 32 finished rays of the calling warp (the tail of its done list), one per lane, at full SIMD width:
// closest-hit / miss + path update exactly as the wavefront shade kernel does (shade_one), then every path that ended
// is replaced by the next primary ray of the pass (one atomic per batch) while there are any. Slots that hold a ray
// again go onto the ready list. Not inlined: the traversal loop keeps its register allocation, and what lives across
// the call is saved around it once per ~30 loop iterations.
// Returns (entries left on the done list) | (entries on the ready list) << 8.
__device__ __noinline__ uint32_t fused_shade_batch(const FusedArgs* f, uint32_t wf_a, uint32_t n_done) {
    using namespace bpt_shade;
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint32_t cnt = min(32u, n_done), first = n_done - cnt;
    uint32_t slot = 0u, sa = 0u, pix = 0u, depth = 0u;
    bool alive = false, need_new = false;
    ShadeOut o;
    o.ro = o.rd = o.st = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < cnt) {
        slot = lds8(wf_a + first + lane);
        sa = wf_a + kFusedSlot0 + slot * 64u;
        const uint4 q2 = lds128(sa + 32u);
        const uint32_t prim = q2.y;
        depth = q2.w >> 8;
        if (prim == kNewPath) need_new = true;
        else {
            const uint4 q0 = lds128(sa), q1 = lds128(sa + 16u), q3 = lds128(sa + 48u);
            pix = q0.w;
            const float4 ro = make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y), __uint_as_float(q0.z), f->p.tmin);
            const float4 rd = make_float4(__uint_as_float(q1.x), __uint_as_float(q1.y), __uint_as_float(q1.z), f->p.tmax);
            const float4 st = make_float4(__uint_as_float(q3.x), __uint_as_float(q3.y), __uint_as_float(q3.z), __uint_as_float(q1.w));
            float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra, rc = ra, rdd = ra;
            if (prim != BPT_MISS) {
                const float4* rp = f->s.srec + 4 * (size_t)(f->s.xforms ? prim % f->s.ntris : prim);
                ra = __ldg(rp); rb = __ldg(rp + 1); rc = __ldg(rp + 2); rdd = __ldg(rp + 3);
            }
            alive = shade_one<false, true>(f->p, f->s, depth, make_uint4(q2.x, 0u, 0u, prim), st, pix, ro, rd, ra, rb, rc, rdd,
                                     f->path_color, nullptr, 0.f, o);
            if (alive) ++depth; else need_new = true;
        }
    }
    const unsigned nm = __ballot_sync(FULL, need_new);
    if (nm) {
        // one atomic per batch: consecutive path ids for the lanes whose path ended. (Handing a warp 256 ids at a time
        // was tried: nothing on the big frames, and small frames — fewer chunks than warps — ran on a fraction of the
        // machine.) The counter runs past npaths while a pass drains: it is 32 bits, a pass has < 2^31 paths.
        const int leader = __ffs(nm) - 1;
        uint32_t base = 0u;
        if ((int)lane == leader) base = atomicAdd(f->path_ctr, (uint32_t)__popc(nm));
        base = __shfl_sync(FULL, base, leader);
        if (need_new) {
            const uint32_t i = base + __popc(nm & lt);
            if (i < f->npaths) {
                uint32_t seed;
                gen_primary(f->p, f->s0, f->npix, i, o.ro, o.rd, seed);
                o.st = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed));
                pix = f->path_base + i;
                depth = 0u;
                alive = true;
            }
        }
    }
    const unsigned am = __ballot_sync(FULL, alive);
    if (alive) {
        const uint32_t oct = (o.rd.x >= 0.f ? 4u : 0u) | (o.rd.y >= 0.f ? 2u : 0u) | (o.rd.z >= 0.f ? 1u : 0u);
        sts128(sa, make_uint4(__float_as_uint(o.ro.x), __float_as_uint(o.ro.y), __float_as_uint(o.ro.z), pix));
        sts128(sa + 16u, make_uint4(__float_as_uint(o.rd.x), __float_as_uint(o.rd.y), __float_as_uint(o.rd.z), __float_as_uint(o.st.w)));
        sts128(sa + 32u, make_uint4(__float_as_uint(safe_rcp(o.rd.x)), __float_as_uint(safe_rcp(o.rd.y)),
                                    __float_as_uint(safe_rcp(o.rd.z)), oct | (depth << 8)));
        sts128(sa + 48u, make_uint4(__float_as_uint(o.st.x), __float_as_uint(o.st.y), __float_as_uint(o.st.z), 0u));
        sts8(wf_a + kFusedReady + __popc(am & lt), slot);
    }
    if (lane == 0u) {  // rays handed to the traversal: the warp's share of bpt_stats.rays
        const uint2 c = lds64(wf_a + kFusedCtl);
        const unsigned long long n = (((unsigned long long)c.y << 32) | c.x) + (unsigned long long)__popc(am);
        sts64(wf_a + kFusedCtl, make_uint2((uint32_t)n, (uint32_t)(n >> 32)));
    }
    __syncwarp();
    return (n_done - cnt) | ((uint32_t)__popc(am) << 8);
}

// TWO_LEVEL (instanced scenes, main.cpp:515-538): nodes = [mesh BVH8 | instance BVH8], records = [triangles | instances];
// an instance record holds the rows of the inverse 3x4 transform. Hitting one pushes what is left of the instance-level
// state plus a sentinel, moves the ray into object space (d is NOT renormalised, so t is the same in both spaces) and
// descends from the mesh root (node 0); popping the sentinel reloads the world-space ray.
//
// FUSED (the path kernel): no ray queue at all. A warp owns kFusedSlots path slots in shared memory; a lane whose ray is
// done leaves {t, primitive} in its slot and takes a slot from the warp's ready list, and when that list is empty the
// warp shades its finished rays 32 at a time (fused_shade_batch), which refills it — with bounce rays and, for every
// path that ended, the next primary ray of the pass. Ray and hit records, path state and the queue compaction of the
// wavefront never touch HBM, and a sample pass is one launch with one tail instead of one per bounce.
template <int BLOCK, int SSTACK, bool STAGED, bool TWO_LEVEL, bool COUNT, bool FUSED>
__global__ void __launch_bounds__(BLOCK, BLOCK == kTraceBlock ? 1 : BPT_TRACE_CTAS_G) k_trace(TraceArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint2* slut = reinterpret_cast<uint2*>(smem_raw + 16);  // byte o of slut[b]: bit p = bit (p ^ o) of b
    unsigned char* spool = smem_raw + 16 + 2048;  // ray pools: 32 rays x 48 B ({o,tmin} {d,tmax} {1/d,octant}) per warp
    unsigned char* sfused = smem_raw + 16 + 2048;  // FUSED instead: FusedArgs, then per warp {done, ready, counters, slots}
    uint2* sstack = reinterpret_cast<uint2*>(smem_raw + trace_smem_fixed(BLOCK, FUSED));
    unsigned char* srecs = smem_raw + trace_smem_fixed(BLOCK, FUSED) + (size_t)SSTACK * BLOCK * sizeof(uint2);  // staged records

    const uint32_t nrays = FUSED ? 0u : *a.count_ptr;
    if (!FUSED && nrays == 0u) return;  // uniform across the grid: an exhausted bounce costs one launch and nothing else

    // ---- STAGED: the whole record array moves into shared memory with TMA bulk copies
    if (STAGED) {
        const uint32_t bytes = a.staged_recs * BPT_REC_BYTES;
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            constexpr uint32_t kChunk = 32768u;
            for (uint32_t off = 0; off < bytes; off += kChunk)
                tma_bulk_g2s(srecs + off, reinterpret_cast<const unsigned char*>(a.recs) + off, min(kChunk, bytes - off), bar);
        }
        mbar_wait(bar, 0);
    }
    for (uint32_t b = threadIdx.x; b < 256u; b += BLOCK) {
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (uint32_t o = 0; o < 8; ++o) {
            uint32_t p = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) p |= ((b >> (k ^ o)) & 1u) << k;
            w[o >> 2] |= p << (8 * (o & 3));
        }
        slut[b] = make_uint2(w[0], w[1]);
    }
    if (FUSED) {
        // arguments of the shade batch into shared memory (the frame index from device memory under graph replay);
        // every slot starts on its warp's done list as "never held a path", so the first batches generate primary rays
        // (one thread, constant indices: a run-time index into the kernel parameters would make the compiler keep a
        // copy of all of them in local memory and read traversal constants from there)
        if (threadIdx.x == 0) *reinterpret_cast<FusedArgs*>(sfused) = a.f;
        unsigned char* wb = sfused + kFusedArgsBytes + (threadIdx.x >> 5) * kFusedWarpBytes;
        for (uint32_t i = threadIdx.x & 31u; i < (uint32_t)kFusedSlots; i += 32u) {
            wb[i] = (unsigned char)i;
            *reinterpret_cast<uint4*>(wb + kFusedSlot0 + i * 64u + 32u) = make_uint4(0u, kNewPath, 0u, 0u);
        }
        if ((threadIdx.x & 31u) == 0u) *reinterpret_cast<uint2*>(wb + kFusedCtl) = make_uint2(0u, 0u);  // rays traced
        __syncthreads();
        if (threadIdx.x == 0 && a.f.frame_dev) reinterpret_cast<FusedArgs*>(sfused)->p.frame = *a.f.frame_dev;
    }
    __syncthreads();

    if (!FUSED && blockIdx.x == 0 && threadIdx.x == 0 && a.stat && a.count_rays) atomicAdd(a.stat + BPT_STAT_RAYS, (unsigned long long)nrays);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const FusedArgs* fargs = reinterpret_cast<const FusedArgs*>(sfused);
    // every shared-memory address below is the one base plus compile-time offsets
    uint32_t smem_a;  // opaque to the compiler: it would otherwise rebuild it (S2R SR_CgaCtaId + LEA) at every use under pressure
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_a) : "r"(smem_u32(smem_raw)));
    constexpr uint32_t kFixed = (uint32_t)trace_smem_fixed(BLOCK, FUSED);
    const uint32_t wf_a = smem_a + 16u + 2048u + kFusedArgsBytes + (threadIdx.x >> 5) * kFusedWarpBytes;  // FUSED: the warp's block
    uint32_t n_done = kFusedSlots, n_ready = 0u;  // FUSED, warp-uniform: lengths of the warp's done and ready lists
    const uint32_t stack_a = smem_a + kFixed + threadIdx.x * 8u;  // entry i of this lane: stack_a + i * BLOCK * 8
    const uint32_t srecs_a = smem_a + kFixed + (uint32_t)SSTACK * BLOCK * 8u, slut_a = smem_a + 16u;
    const uint32_t pool_a = smem_a + 16u + 2048u + (threadIdx.x >> 5) * 1536u;
    uint2 lstack[kLocalStack];
    volatile uint32_t park[FUSED ? 20 : 1];  // FUSED: a lane's traversal state while the warp shades
    const uint32_t magic = a.magic;
    const int refill_below = a.refill_below, steps_per_refill = a.steps_per_refill;
    const int tris_per_step = STAGED ? a.staged_tris_per_step : 1;

    RayState r;
    if (FUSED) r.tmin = a.f.p.tmin;  // every ray of a frame has the frame's tmin / tmax
    uint2 G = make_uint2(0u, 0u);  // node group: x = first internal child, y = hits by priority << 24 | internal mask
    uint2 T = make_uint2(0u, 0u);  // triangle group: x = node the triangles belong to, y = hit bits (valid layout)
    uint32_t Tb = 0u, Tv = 0u;     // tri_base and valid word of node T.x
    uint32_t octsel = 0u;          // byte-permute selector that picks byte `oct` of a slut row into byte 3
    uint32_t inst_base = 0u;       // TWO_LEVEL: instance * mesh triangles while inside an instance
    int sp = 0;
    uint32_t ray_idx = FUSED ? 0xffu : 0u;  // FUSED: the lane's path slot (0xff: none)
    bool active = false, exhausted = false;
    uint32_t pool_base = 0u, pool_count = 0u, pool_next = 0u;  // warp-uniform: the warp's ray pool in shared memory
    unsigned long long cnt_nodes = 0, cnt_tris = 0, cnt_witer = 0, cnt_wnode = 0, cnt_wtri = 0, cnt_liter = 0;

#define BPT_PUSH(E)                                                                               \
    {                                                                                             \
        if (sp < SSTACK) sts64(stack_a + sp * (BLOCK * 8), (E)); else lstack[sp - SSTACK] = (E);  \
        ++sp;                                                                                     \
    }
#define BPT_LEADER() ((__activemask() & lt) == 0u)

    for (;;) {
        unsigned actmask = __ballot_sync(FULL, active);
        if (FUSED) {
            if (__popc(actmask) < refill_below) {
                // finished rays go onto the done list; idle lanes take ready slots; an empty ready list is refilled by
                // shading the tail of the done list. With all 64 slots alive the done list holds >= 32 entries whenever
                // the ready list runs dry, so the batches are full until the pass drains.
                const bool fin = !active && ray_idx != 0xffu;
                const unsigned fm = __ballot_sync(FULL, fin);
                if (fin) {
                    sts8(wf_a + n_done + __popc(fm & lt), ray_idx);
                    // the shading record of the hit on its way into the L2 while the ray waits for its batch
                    const uint32_t hp = lds32(wf_a + kFusedSlot0 + ray_idx * 64u + 36u);
                    if (!TWO_LEVEL && hp != BPT_MISS) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.f.s.srec + 4 * (size_t)hp));
                    ray_idx = 0xffu;
                }
                n_done += __popc(fm);
                __syncwarp();
                unsigned idle = ~actmask;
                for (;;) {
                    if (n_ready == 0u) {
                        if (n_done == 0u) break;
                        // The lanes' traversal state waits in local memory while the warp shades: the batch needs the
                        // registers, and spelled out like this the allocator does not keep loop state spilled instead.
                        // One-level scenes reload the ray itself from the lane's path slot.
                        if (TWO_LEVEL) {
                            park[9] = __float_as_uint(r.ox); park[10] = __float_as_uint(r.oy); park[11] = __float_as_uint(r.oz);
                            park[12] = __float_as_uint(r.dx); park[13] = __float_as_uint(r.dy); park[14] = __float_as_uint(r.dz);
                            park[15] = __float_as_uint(r.idx); park[16] = __float_as_uint(r.idy); park[17] = __float_as_uint(r.idz);
                            park[18] = r.oct; park[19] = inst_base;
                        }
                        park[0] = __float_as_uint(r.tbest); park[1] = r.hprim;
                        park[2] = G.x; park[3] = G.y; park[4] = T.x; park[5] = T.y; park[6] = Tb; park[7] = Tv;
                        park[8] = (uint32_t)sp | (ray_idx << 8) | (active ? 0x10000u : 0u);
                        uint32_t res = fused_shade_batch(fargs, wf_a, n_done);
                        res = __shfl_sync(FULL, res, 0);  // warp-uniform, and known to be
                        n_done = res & 0xffu; n_ready = res >> 8;
                        r.tbest = __uint_as_float(park[0]); r.hprim = park[1];
                        G.x = park[2]; G.y = park[3]; T.x = park[4]; T.y = park[5]; Tb = park[6]; Tv = park[7];
                        { const uint32_t pk = park[8]; sp = (int)(pk & 0xffu); ray_idx = (pk >> 8) & 0xffu; active = (pk & 0x10000u) != 0u; }
                        if (TWO_LEVEL) {
                            r.ox = __uint_as_float(park[9]); r.oy = __uint_as_float(park[10]); r.oz = __uint_as_float(park[11]);
                            r.dx = __uint_as_float(park[12]); r.dy = __uint_as_float(park[13]); r.dz = __uint_as_float(park[14]);
                            r.idx = __uint_as_float(park[15]); r.idy = __uint_as_float(park[16]); r.idz = __uint_as_float(park[17]);
                            r.oct = park[18]; inst_base = park[19];
                        } else {  // every lane, live or not (an idle lane reads some slot: nothing looks at its ray)
                            const uint32_t sa = wf_a + kFusedSlot0 + (ray_idx & (kFusedSlots - 1u)) * 64u;
                            const uint4 q0 = lds128(sa), q1 = lds128(sa + 16u), q2 = lds128(sa + 32u);
                            r.ox = __uint_as_float(q0.x); r.oy = __uint_as_float(q0.y); r.oz = __uint_as_float(q0.z);
                            r.dx = __uint_as_float(q1.x); r.dy = __uint_as_float(q1.y); r.dz = __uint_as_float(q1.z);
                            r.idx = __uint_as_float(q2.x); r.idy = __uint_as_float(q2.y); r.idz = __uint_as_float(q2.z);
                            r.oct = q2.w & 7u;
                        }
                        octsel = r.oct << 12;
                        continue;  // an empty ready list again: every shaded path ended and the pass has no primary ray left
                    }
                    const uint32_t rank = __popc(idle & lt);
                    if (!active && rank < n_ready) {
                        ray_idx = lds8(wf_a + kFusedReady + (n_ready - 1u - rank));
                        const uint32_t sa = wf_a + kFusedSlot0 + ray_idx * 64u;
                        const uint4 q0 = lds128(sa), q1 = lds128(sa + 16u), q2 = lds128(sa + 32u);
                        r.ox = __uint_as_float(q0.x); r.oy = __uint_as_float(q0.y); r.oz = __uint_as_float(q0.z);
                        r.dx = __uint_as_float(q1.x); r.dy = __uint_as_float(q1.y); r.dz = __uint_as_float(q1.z); r.tbest = a.f.p.tmax;
                        r.idx = __uint_as_float(q2.x); r.idy = __uint_as_float(q2.y); r.idz = __uint_as_float(q2.z);
                        r.oct = q2.w & 7u;
                        octsel = r.oct << 12;
                        r.hprim = BPT_MISS;
                        inst_base = 0u;
                        G = make_uint2(a.root, 0x80000000u);
                        T = make_uint2(0u, 0u);
                        sp = 0;
                        active = true;
                    }
                    n_ready -= min(n_ready, (uint32_t)__popc(idle));
                    __syncwarp();  // the list and slot reads are done before a batch rewrites them
                    idle = ~__ballot_sync(FULL, active);
                    if (!idle) break;
                }
                actmask = ~idle;
            }
        } else
        if (__popc(actmask) < refill_below && (!exhausted || pool_next < pool_count)) {
            // Refill idle lanes from the warp's ray pool: 32 consecutive rays fetched with one atomic and two coalesced
            // 128-bit loads per lane into shared memory; lanes that finish take the next pool entries without touching
            // global memory, so the warp can top itself up every iteration and runs ~31 live lanes instead of ~26.
            unsigned idle = ~actmask;
            while (idle) {
                if (pool_next >= pool_count) {
                    if (exhausted) break;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(a.fetch_ctr, 32u);
                    base = __shfl_sync(FULL, base, 0);
                    if (base >= nrays) { exhausted = true; break; }
                    pool_base = base; pool_count = min(32u, nrays - base); pool_next = 0u;
                    if (lane < pool_count) {
                        // all 32 lanes set their pool ray up (1/d, octant) here, at full SIMD width, so that a lane
                        // that takes a ray later only copies 48 bytes
                        const float4 ro = __ldg(&a.rays[2 * (size_t)(base + lane)]);
                        const float4 rd = __ldg(&a.rays[2 * (size_t)(base + lane) + 1]);
                        const uint32_t oct = (rd.x >= 0.f ? 4u : 0u) | (rd.y >= 0.f ? 2u : 0u) | (rd.z >= 0.f ? 1u : 0u);
                        sts128f(pool_a + lane * 48u, ro);
                        sts128f(pool_a + lane * 48u + 16u, rd);
                        sts128f(pool_a + lane * 48u + 32u, make_float4(safe_rcp(rd.x), safe_rcp(rd.y), safe_rcp(rd.z), __uint_as_float(oct)));
                    }
                    __syncwarp();
                }
                const uint32_t avail = pool_count - pool_next;
                const uint32_t rank = __popc(idle & lt);
                if (!active && rank < avail) {
                    const uint32_t slot = pool_next + rank;
                    const float4 ro = as_float4(lds128(pool_a + slot * 48u)), rd = as_float4(lds128(pool_a + slot * 48u + 16u));
                    const float4 ri = as_float4(lds128(pool_a + slot * 48u + 32u));
                    ray_idx = pool_base + slot;
                    r.ox = ro.x; r.oy = ro.y; r.oz = ro.z; r.tmin = ro.w;
                    r.dx = rd.x; r.dy = rd.y; r.dz = rd.z; r.tbest = rd.w;
                    r.idx = ri.x; r.idy = ri.y; r.idz = ri.z;
                    r.oct = __float_as_uint(ri.w);
                    octsel = r.oct << 12;
                    r.hprim = BPT_MISS;
                    inst_base = 0u;
                    G = make_uint2(a.root, 0x80000000u);  // root through priority bit 31, internal mask 0
                    T = make_uint2(0u, 0u);
                    sp = 0;
                    active = true;
                }
                pool_next += min(avail, (uint32_t)__popc(idle));
                __syncwarp();  // pool reads are done before a refill of the pool overwrites it
                idle = ~__ballot_sync(FULL, active);
            }
            actmask = ~idle;
        }
        if (actmask == 0u) break;

        // Every lane of the warp runs the loop below, live or not (an idle lane has no node bits, no triangle bits and
        // an empty stack), and the warp reconverges with __syncwarp() in front of every phase.
#pragma unroll 1
        for (int it = 0; it < steps_per_refill; ++it) {
            if (COUNT) {
                const unsigned m = __ballot_sync(FULL, active);
                if (lane == 0 && m) ++cnt_witer;
                if (active) ++cnt_liter;
            }
#if BPT_NODE_STEPS > 1
#pragma unroll
            for (int round = 0; round < BPT_NODE_STEPS; ++round) {
#endif
            // ---------------- pop when out of node work: a node group always, a parked triangle group only when
            //                  none is draining (it stays on the stack until then)
            if (active && !(G.y & 0xff000000u) && sp > 0) {
                const uint2 e = sp <= SSTACK ? lds64(stack_a + (sp - 1) * (BLOCK * 8)) : lstack[sp - 1 - SSTACK];
                if (e.y & 0xff000000u) { G = e; --sp; }
                else if (TWO_LEVEL && e.y == 0u) {
                    if (T.y == 0u) {  // sentinel: the instance is done (its triangles too) -> back to world space
                        --sp;
                        float4 ro, rd;
                        if (FUSED) {
                            const uint32_t sa = wf_a + kFusedSlot0 + ray_idx * 64u;
                            ro = as_float4(lds128(sa)); rd = as_float4(lds128(sa + 16u));
                        } else {
                            ro = __ldg(&a.rays[2 * (size_t)ray_idx]);
                            rd = __ldg(&a.rays[2 * (size_t)ray_idx + 1]);
                        }
                        r.ox = ro.x; r.oy = ro.y; r.oz = ro.z;
                        r.dx = rd.x; r.dy = rd.y; r.dz = rd.z;
                        r.idx = safe_rcp(rd.x);
                        r.idy = safe_rcp(rd.y);
                        r.idz = safe_rcp(rd.z);
                        r.oct = (rd.x >= 0.f ? 4u : 0u) | (rd.y >= 0.f ? 2u : 0u) | (rd.z >= 0.f ? 1u : 0u);
                        octsel = r.oct << 12;
                        inst_base = 0u;
                    }
                }
                else if (T.y == 0u) {
                    T = e; --sp;
                    if (STAGED) { const uint2 h = lds64(srecs_a + e.x * BPT_REC_BYTES + 8u); Tv = h.x; Tb = h.y; }
                    else {
                        const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(a.recs) + (size_t)e.x * BPT_REC_BYTES + 8));
                        Tv = h.x; Tb = h.y;
                    }
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
                U8 v0, v1;
                if (STAGED) {  // staged: the whole BVH, or the top of the tree
                    const uint32_t np = srecs_a + node * BPT_REC_BYTES;
                    v0 = lds256(np); v1 = lds256(np + 32u);
                } else {
                    const unsigned char* np = reinterpret_cast<const unsigned char*>(a.recs) + (size_t)node * BPT_REC_BYTES;
                    v0 = ldg256(np); v1 = ldg256(np + 32);
                }
                // v0: org_lo org_hi e|valid child_base | qlox qlox qloy qloy   v1: qloz qloz qhix qhix | qhiy qhiy qhiz qhiz
                const int lvl = TWO_LEVEL ? (node >= a.root ? 1 : 0) : 0;  // instance-level nodes live behind the mesh's
                // origin: 21-bit grid coordinate c -> float 2^23 + c by OR-ing the exponent pattern (one LOP3), the 2^23
                // bias is folded into the grid offset on the host (gbias = glo - 2^23 * gstep); the builder evaluates
                // the same expression, so both sides agree on the origin bit for bit
                const float pox = fmaf(__uint_as_float((v0.lo.x & 0x1fffffu) | 0x4b000000u), a.gstep[lvl][0], a.gbias[lvl][0]);
                const float poy = fmaf(__uint_as_float((__funnelshift_r(v0.lo.x, v0.lo.y, 21) & 0x1fffffu) | 0x4b000000u), a.gstep[lvl][1], a.gbias[lvl][1]);
                const float poz = fmaf(__uint_as_float(((v0.lo.y >> 10) & 0x1fffffu) | 0x4b000000u), a.gstep[lvl][2], a.gbias[lvl][2]);
                const float step = __uint_as_float((v0.lo.z >> 1) & 0x7f800000u);  // exponent byte on top of the word
                const float adx = step * r.idx, ady = step * r.idy, adz = step * r.idz;
                const float bx = fmaf(adx, -kByteBias, (pox - r.ox) * r.idx);
                const float by = fmaf(ady, -kByteBias, (poy - r.oy) * r.idy);
                const float bz = fmaf(adz, -kByteBias, (poz - r.oz) * r.idz);
                // near / far plane bytes by direction sign
                const bool nx = r.idx < 0.f, ny = r.idy < 0.f, nz = r.idz < 0.f;
                const uint32_t xn0 = nx ? v1.lo.z : v0.hi.x, xn1 = nx ? v1.lo.w : v0.hi.y, xf0 = nx ? v0.hi.x : v1.lo.z, xf1 = nx ? v0.hi.y : v1.lo.w;
                const uint32_t yn0 = ny ? v1.hi.x : v0.hi.z, yn1 = ny ? v1.hi.y : v0.hi.w, yf0 = ny ? v0.hi.z : v1.hi.x, yf1 = ny ? v0.hi.w : v1.hi.y;
                const uint32_t zn0 = nz ? v1.hi.z : v1.lo.x, zn1 = nz ? v1.hi.w : v1.lo.y, zf0 = nz ? v1.lo.x : v1.hi.z, zf1 = nz ? v1.lo.y : v1.hi.w;
                uint32_t hit = test_quad<0>(xn0, yn0, zn0, xf0, yf0, zf0, adx, ady, adz, bx, by, bz, r.tmin, r.tbest, magic);
                hit |= test_quad<1>(xn1, yn1, zn1, xf1, yf1, zf1, adx, ady, adz, bx, by, bz, r.tmin, r.tbest, magic);
                const uint32_t valid = v0.lo.z;  // bits 0..15 triangle counts, 16..23 internal mask (24..31: exponent)
                hit &= valid;
                // internal hits in visiting priority (bit p <- slot p ^ oct) into byte 3, internal mask into byte 0;
                // bytes 1 and 2 of a node group are never looked at
                const uint2 row = lds64(slut_a + ((hit >> 16) << 3));
                const uint32_t prio = __byte_perm(row.x, row.y, octsel);  // byte 3 = row byte `oct`
                G.x = v0.lo.w;
                G.y = __byte_perm(valid, prio, 0x7002);
                if (hit & 0xffffu) {
                    if (T.y) BPT_PUSH(T)  // a group is still draining: park it, it is popped like any other entry
                    T.x = node;
                    T.y = hit & 0xffffu;  // per hit leaf slot: how many of its triangles are still to be tested
                    Tb = v0.lo.w; Tv = valid;
                }
            }
#if BPT_NODE_STEPS > 1
            }
#endif
            // ---------------- triangle step: one triangle of the lane's pending group
            __syncwarp();
#pragma unroll 1
            for (int q = 0; q < tris_per_step; ++q)  // global instance: exactly one
            if (T.y) {
                if (COUNT) { if (BPT_LEADER()) ++cnt_wtri; ++cnt_tris; }
                // highest pending slot s, its last untested triangle k; the record sits behind the node's internal
                // children and the triangles of the lower leaf slots
                const uint32_t sh = (31u - __clz(T.y)) & ~1u;            // 2 * s
                const uint32_t k = ((T.y >> sh) & 3u) - 1u;
                T.y -= 1u << sh;
                const uint32_t below = Tv & ~(0xffffffffu << sh);        // counts of the slots below s
                const uint32_t tri = Tb + __popc(Tv & 0xff0000u) + __popc(below & 0x5555u) + 2u * __popc(below & 0xaaaau) + k;
                U8 w0, w1;  // w0: ru rv   w1: rw | prim - - -
                if (STAGED) {
                    const uint32_t tp = srecs_a + tri * BPT_REC_BYTES;
                    w0 = lds256(tp); w1.lo = lds128(tp + 32u); w1.hi = make_uint4(lds32(tp + 48u), 0u, 0u, 0u);
                } else {
                    const unsigned char* tp = reinterpret_cast<const unsigned char*>(a.recs) + (size_t)tri * BPT_REC_BYTES;
                    w0 = ldg256(tp); w1 = ldg256(tp + 32);
                }
                if (TWO_LEVEL && tri >= a.root) {
                    // instance record: w0 = rows 0,1 and w1.lo = row 2 of the inverse transform, w1.hi.x = instance
                    if (G.y & 0xff000000u) BPT_PUSH(G)
                    if (T.y) BPT_PUSH(T)
                    BPT_PUSH(make_uint2(0u, 0u))
                    const float m00 = __uint_as_float(w0.lo.x), m01 = __uint_as_float(w0.lo.y), m02 = __uint_as_float(w0.lo.z), m03 = __uint_as_float(w0.lo.w);
                    const float m10 = __uint_as_float(w0.hi.x), m11 = __uint_as_float(w0.hi.y), m12 = __uint_as_float(w0.hi.z), m13 = __uint_as_float(w0.hi.w);
                    const float m20 = __uint_as_float(w1.lo.x), m21 = __uint_as_float(w1.lo.y), m22 = __uint_as_float(w1.lo.z), m23 = __uint_as_float(w1.lo.w);
                    const float ox = dot3p(m00, m01, m02, r.ox, r.oy, r.oz, m03), oy = dot3p(m10, m11, m12, r.ox, r.oy, r.oz, m13),
                                oz = dot3p(m20, m21, m22, r.ox, r.oy, r.oz, m23);
                    const float dx = dot3(m00, m01, m02, r.dx, r.dy, r.dz), dy = dot3(m10, m11, m12, r.dx, r.dy, r.dz),
                                dz = dot3(m20, m21, m22, r.dx, r.dy, r.dz);
                    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
                    r.idx = safe_rcp(dx);
                    r.idy = safe_rcp(dy);
                    r.idz = safe_rcp(dz);
                    r.oct = (dx >= 0.f ? 4u : 0u) | (dy >= 0.f ? 2u : 0u) | (dz >= 0.f ? 1u : 0u);
                    octsel = r.oct << 12;
                    inst_base = w1.hi.x * a.num_mesh_tris;
                    G = make_uint2(0u, 0x80000000u);  // mesh root
                    T = make_uint2(0u, 0u);
                } else {
                const float rwx = __uint_as_float(w1.lo.x), rwy = __uint_as_float(w1.lo.y), rwz = __uint_as_float(w1.lo.z);
                const float oz = dot3p(r.ox, r.oy, r.oz, rwx, rwy, rwz, __uint_as_float(w1.lo.w));
                const float dz = dot3(r.dx, r.dy, r.dz, rwx, rwy, rwz);
                float rdz;  // MUFU.RCP alone: a denormal dz (ray in the triangle's plane) gives inf / NaN, which fails the range test
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rdz) : "f"(dz));
                const float t = -oz * rdz;
#if BPT_TRI_LAZY_UV
                if (t >= r.tmin && t <= r.tbest)
#endif
                {
                    // u and v are evaluated for every tested triangle, not only behind the distance test: ptxas sinks the
                    // load of the record's first half into that branch otherwise, and the warp pays a third exposed
                    // memory round trip per iteration (10 % of all stall samples, profiles/r2b_*) for 1 lane in 6 it
                    // spares the arithmetic
                    const float rux = __uint_as_float(w0.lo.x), ruy = __uint_as_float(w0.lo.y), ruz = __uint_as_float(w0.lo.z);
                    const float rvx = __uint_as_float(w0.hi.x), rvy = __uint_as_float(w0.hi.y), rvz = __uint_as_float(w0.hi.z);
                    const float ou = dot3p(r.ox, r.oy, r.oz, rux, ruy, ruz, __uint_as_float(w0.lo.w));
                    const float du = dot3(r.dx, r.dy, r.dz, rux, ruy, ruz);
                    const float u = fmaf(t, du, ou);
                    const float ov = dot3p(r.ox, r.oy, r.oz, rvx, rvy, rvz, __uint_as_float(w0.hi.w));
                    const float dv = dot3(r.dx, r.dy, r.dz, rvx, rvy, rvz);
                    const float v = fmaf(t, dv, ov);
                    // equal distance (exact duplicate triangles): lowest primitive id wins
                    const uint32_t prim = inst_base + w1.hi.x;
                    if (t >= r.tmin && t <= r.tbest && u >= 0.f && v >= 0.f && u + v <= 1.f && (t < r.tbest || prim < r.hprim)) {
                        r.tbest = t;
                        r.hprim = prim;
                    }
                }
                }
            }
            // ---------------- terminate
            __syncwarp();
            if (active && !(G.y & 0xff000000u) && T.y == 0u && sp == 0) {
                // {t, -, -, prim}: u, v are re-derived from the original vertices by the consumer (shade.cu)
                if (FUSED) sts64(wf_a + kFusedSlot0 + ray_idx * 64u + 32u, make_uint2(__float_as_uint(r.tbest), r.hprim));
                else a.hits[ray_idx] = make_uint4(__float_as_uint(r.tbest), 0u, 0u, r.hprim);
                active = false;
            }
        }
    }
#undef BPT_PUSH
#undef BPT_LEADER
    if (FUSED && lane == 0u && a.stat && a.count_rays) {
        const uint2 c = lds64(wf_a + kFusedCtl);
        atomicAdd(a.stat + BPT_STAT_RAYS, ((unsigned long long)c.y << 32) | c.x);
    }
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

// shared memory the traversal kernel needs with `staged_recs` records staged
size_t trace_smem_bytes(uint32_t staged_recs, int block, bool fused) {
    return trace_smem_fixed(block, fused) + (size_t)kTraceSmemStack * block * sizeof(uint2) + (size_t)staged_recs * BPT_REC_BYTES;
}

cudaError_t trace_configure() {
    cudaError_t e;
#define CFG(B, S, L, C, F)                                                                            \
    if ((e = cudaFuncSetAttribute(k_trace<B, kTraceSmemStack, S, L, C, F>,                             \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, B == kTraceBlock ? kTraceMaxSmem : kTraceMaxSmem / BPT_TRACE_CTAS_G)) != cudaSuccess) \
        return e;
#define CFG2(B, S, L, C) CFG(B, S, L, C, false) CFG(B, S, L, C, true)
    CFG2(BPT_TRACE_BLOCK_G, false, false, false) CFG2(BPT_TRACE_BLOCK_G, false, false, true) CFG2(kTraceBlock, true, false, false)
    CFG2(kTraceBlock, true, false, true) CFG2(BPT_TRACE_BLOCK_G, false, true, false) CFG2(BPT_TRACE_BLOCK_G, false, true, true)
    CFG2(kTraceBlock, true, true, false) CFG2(kTraceBlock, true, true, true)
#undef CFG2
#undef CFG
    return cudaSuccess;
}

void trace_launch(const TraceArgs& a, unsigned num_sms, bool staged, bool two_level, bool count, bool fused, cudaStream_t st) {
    const int block = staged ? kTraceBlock : BPT_TRACE_BLOCK_G;
    const size_t smem = trace_smem_bytes(a.staged_recs, block, fused);
    const unsigned grid = staged ? num_sms : num_sms * (unsigned)BPT_TRACE_CTAS_G;
#define GO(B, S, L, C, F) k_trace<B, kTraceSmemStack, S, L, C, F><<<grid, B, smem, st>>>(a)
#define GO2(B, S, L, C) { if (fused) GO(B, S, L, C, true); else GO(B, S, L, C, false); }
    if (staged) {
        if (two_level) { if (count) GO2(kTraceBlock, true, true, true) else GO2(kTraceBlock, true, true, false) }
        else { if (count) GO2(kTraceBlock, true, false, true) else GO2(kTraceBlock, true, false, false) }
    } else {
        if (two_level) { if (count) GO2(BPT_TRACE_BLOCK_G, false, true, true) else GO2(BPT_TRACE_BLOCK_G, false, true, false) }
        else { if (count) GO2(BPT_TRACE_BLOCK_G, false, false, true) else GO2(BPT_TRACE_BLOCK_G, false, false, false) }
    }
#undef GO2
#undef GO
}
