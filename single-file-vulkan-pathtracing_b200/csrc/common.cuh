// common.cuh — shared device/host declarations of libbpt (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bpt.h"

#define BPT_MISS 0xffffffffu
#define BPT_NUM_SMS_DEFAULT 148

// ------------------------------------------------------------------ acceleration-structure records (64 bytes)
// The traversal kernel is bound by the L1 data pipe: every 32-byte sector a lane receives costs one cycle of it
// (every lane walks its own node, nothing coalesces), so the records are sized in sectors. One array of 64-byte,
// 64-byte-aligned records holds BVH8 nodes AND triangle (or instance) records: two 256-bit loads (LDG.E.256) each.
//
// Node8 — compressed 8-wide node after Ylitie, Karras, Laine 2017, squeezed from their 80 bytes to 64:
//   w0,w1 : grid origin, 3 x 21-bit unsigned coordinates c on the scene grid: x = bits 0..20, y = 21..41,
//           z = 42..62; origin = fmaf(float(2^23 + c), grid_step, grid_bias) per axis, grid_step / grid_bias are
//           per-BVH kernel arguments (grid_bias = grid_lo - 2^23 * grid_step); builder and kernel evaluate the same
//           float expression, so they agree bit for bit
//   w2    : bits 0..15 = triangle count (0..3) of leaf child s in bits 2s..2s+1; bits 16..23 = internal-child mask;
//           bits 24..31 = biased exponent e of the child-box grid step 2^(e-127), shared by the three axes
//   w3    : child_base = record index of the node's first child record. Children records are contiguous: first the
//           internal children (nodes) in slot order, then the triangles of the leaf children in (slot, k) order —
//           one base pointer serves both, which is what frees the bytes for the planes
//   w4..15: qlo_x[8] qlo_y[8] | qlo_z[8] qhi_x[8] qhi_y[8] qhi_z[8]   (first sector ends after qlo_y)
// Child boxes are origin + q * 2^(e-127) per axis; slot bit 2/1/0 = child on the +x/+y/+z side.
struct __align__(32) Node8 {
    uint32_t org_lo, org_hi;
    uint32_t e_valid;
    uint32_t child_base;
    uint8_t qlox[8], qloy[8];
    uint8_t qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 64, "Node8 must be 64 bytes");

// Woop-transformed triangle: three rows (r.xyz, r.w) of the affine map that sends
// v0,v1,v2,v0+n to (0,0,0),(1,0,0),(0,1,0),(0,0,1): row 0 -> u, row 1 -> v, row 2 -> w; plus the primitive id.
// First sector = rows w and u... see the order below: the kernel reads {rw, prim} first (plane distance test) and
// the second sector {ru, rv} only for triangles whose plane the ray segment reaches.
// Instance records of two-level scenes reuse the struct: rows of the inverse 3x4 transform + instance id.
struct __align__(32) WoopTri {
    float4 ru, rv, rw;
    uint32_t prim, pad0, pad1, pad2;
};
static_assert(sizeof(WoopTri) == 64, "WoopTri must be 64 bytes");
#define BPT_REC_BYTES 64u
#define BPT_GRID_BITS 21

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
