// common.cuh — shared device/host declarations of libbpt (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bpt.h"

#define BPT_MISS 0xffffffffu
#define BPT_NUM_SMS_DEFAULT 148

// ------------------------------------------------------------------ BVH8 node (96 bytes)
// Compressed 8-wide node after Ylitie, Karras, Laine 2017, re-laid out for sm_100: three 32-byte words, 32-byte
// aligned, so a lane fetches a node with three 256-bit loads (LDG.E.256). The traversal kernel is bound by L1
// wavefronts — every lane walks its own node, so each load instruction costs one wavefront per lane whatever its
// width — and 3 x 32 B costs 40 % fewer wavefronts than the 5 x 16 B of the classic 80-byte node. The per-child
// meta bytes are replaced by one validity word so that the kernel spends one predicated OR per child (trace.cu).
//   v0: px, py, pz (grid origin, f32) | sx, sy, sz (grid step per axis, f32, a power of two) | child_base | tri_base
//   v1: valid | 0 | qlo_x[0..7] | qlo_y[0..7] | qlo_z[0..7]
//   v2: qhi_x[0..7] | qhi_y[0..7] | qhi_z[0..7] | 0 | 0
// valid: bit 24+s set   = child slot s is an internal node; the internal children of a node are
//                         consecutive nodes from child_base, in slot order
//        bits 3s..3s+2  = unary triangle count (001, 011, 111) of leaf child s, 0 if s is internal
//                         or empty; the triangles of a node are consecutive leaf slots from tri_base
//                         in (slot, k) order, so bit b of the low 24 is triangle
//                         tri_base + popc(valid & ((1 << b) - 1) & 0xffffff)
// Child boxes are origin + q * step per axis; slot bit 2/1/0 = child on the +x/+y/+z side.
struct __align__(32) Node8 {
    float px, py, pz;
    float sx, sy, sz;
    uint32_t child_base, tri_base;
    uint32_t valid, pad0;
    uint8_t qlox[8], qloy[8], qloz[8];
    uint8_t qhix[8], qhiy[8], qhiz[8];
    uint32_t pad1, pad2;
};
static_assert(sizeof(Node8) == 96, "Node8 must be 96 bytes");

// Woop-transformed triangle: three rows (r.xyz, r.w) of the affine map that sends
// v0,v1,v2,v0+n to (0,0,0),(1,0,0),(0,1,0),(0,0,1): row 0 -> u, row 1 -> v, row 2 -> w; plus the primitive id
// the leaf slot holds. 64 bytes, 32-byte aligned: two 256-bit loads per test, and the hit record / the
// duplicate-triangle tie-break get the primitive id without a second gather.
struct __align__(32) WoopTri {
    float4 ru, rv, rw;
    uint32_t prim, pad0, pad1, pad2;
};
static_assert(sizeof(WoopTri) == 64, "WoopTri must be 64 bytes");
#define BPT_NODE_BYTES 96u
#define BPT_TRI_BYTES 64u

// ------------------------------------------------------------------ wavefront records (SoA)
// ray   : 2 x float4  {ox,oy,oz,tmin} {dx,dy,dz,tmax}
// hit   : 1 x uint4   {t,u,v (float bits), prim}
// state : 1 x float4  {w.r,w.g,w.b, seed bits}  + 1 x uint32 tile-local pixel index

// ------------------------------------------------------------------ RNG (shaders/common.glsl:13-37)
__host__ __device__ __forceinline__ uint32_t bpt_pcg(uint32_t& state) {
    uint32_t prev = state * 747796405u + 2891336453u;
    uint32_t word = ((prev >> ((prev >> 28u) + 4u)) ^ prev) * 277803737u;
    state = prev;
    return (word >> 22u) ^ word;
}
__host__ __device__ __forceinline__ void bpt_pcg2d(uint32_t& x, uint32_t& y) {
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
#ifdef __CUDACC__
// float(val) * (1.0 / float(0xffffffffu)): the divisor rounds to 2^32, so the scale is 2^-32
// and the result can be exactly 1.0 (common.glsl:33-37).
__device__ __forceinline__ float bpt_rand(uint32_t& seed) {
    return __fmul_rn(__uint2float_rn(bpt_pcg(seed)), 2.3283064365386963e-10f);
}
#endif

// ------------------------------------------------------------------ kernel launchers (host side)
struct BuildBuffers;  // build.cu
struct bpt_context;

// error plumbing
#define BPT_CUDA_TRY(ctx, expr)                                                         \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) return bpt_fail_cuda((ctx), _e, #expr, __FILE__, __LINE__); \
    } while (0)
int bpt_fail_cuda(bpt_context* ctx, cudaError_t e, const char* what, const char* file, int line);
int bpt_fail(bpt_context* ctx, int code, const char* fmt, ...);
