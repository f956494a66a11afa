// common.cuh — shared device/host declarations of libbpt (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bpt.h"

#define BPT_MISS 0xffffffffu
#define BPT_NUM_SMS_DEFAULT 148

// ------------------------------------------------------------------ BVH8 node (80 bytes)
// Compressed 8-wide node after Ylitie, Karras, Laine 2017, with the per-child meta bytes replaced by
// one validity word so that the traversal kernel spends one predicated OR per child (trace.cu).
// Five 16-byte words: a lane fetches a node with five 128-bit loads, and a TMA bulk copy moves whole
// blocks of nodes into shared memory.
//   w0: px, py, pz (grid origin, f32) | ex, ey, ez (u8 biased exponents of the grid step), 0
//   w1: child_base (u32), tri_base (u32), valid (u32), 0
//   w2: qlo_x[0..7], qlo_y[0..7]
//   w3: qlo_z[0..7], qhi_x[0..7]
//   w4: qhi_y[0..7], qhi_z[0..7]
// valid: bit 24+s set   = child slot s is an internal node; the internal children of a node are
//                         consecutive nodes from child_base, in slot order
//        bits 3s..3s+2  = unary triangle count (001, 011, 111) of leaf child s, 0 if s is internal
//                         or empty; the triangles of a node are consecutive leaf slots from tri_base
//                         in (slot, k) order, so bit b of the low 24 is triangle
//                         tri_base + popc(valid & ((1 << b) - 1) & 0xffffff)
// Child boxes are origin + q * 2^(e-127) per axis; slot bit 2/1/0 = child on the +x/+y/+z side.
struct __align__(16) Node8 {
    float px, py, pz;
    uint8_t ex, ey, ez, pad0;
    uint32_t child_base, tri_base;
    uint32_t valid, pad1;
    uint8_t qlox[8], qloy[8];
    uint8_t qloz[8], qhix[8];
    uint8_t qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

// Woop-transformed triangle: three rows (r.xyz, r.w) of the affine map that sends
// v0,v1,v2,v0+n to (0,0,0),(1,0,0),(0,1,0),(0,0,1): row 0 -> u, row 1 -> v, row 2 -> w.
struct __align__(16) WoopTri {
    float4 ru, rv, rw;
};
static_assert(sizeof(WoopTri) == 48, "WoopTri must be 48 bytes");

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
