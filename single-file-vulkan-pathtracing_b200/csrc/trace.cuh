// trace.cuh — launch interface of the traversal kernel (trace.cu).
#pragma once
#include "common.cuh"

constexpr int kTraceBlock = 1024;      // one persistent CTA of 32 warps per SM
constexpr int kTraceSmemStack = 8;     // per-lane stack entries held in shared memory
constexpr int kTraceMaxSmem = 227 * 1024;

struct TraceArgs {
    const float4* rays;            // 2 per ray: {o, tmin} {d, tmax}
    uint4* hits;                   // {t, u, v, prim}
    const uint32_t* count_ptr;     // number of rays (device)
    uint32_t* fetch_ctr;           // global ray fetch counter (zeroed before launch)
    const uint4* nodes;            // BVH8 nodes, 5 x uint4 each, BFS order
    const float4* woop;            // 3 x float4 per triangle, leaf order
    const uint32_t* prim_index;    // leaf slot -> primitive id
    uint32_t top_nodes;            // BFS prefix staged into shared memory
    uint32_t top_tris;             // leading triangles staged into shared memory
    unsigned long long* stat_rays; // += ray count (may be null)
    unsigned long long* stat_nodes;
    unsigned long long* stat_tris;
};

size_t trace_smem_bytes(uint32_t top_nodes, uint32_t top_tris);
cudaError_t trace_configure();
void trace_launch(const TraceArgs& a, unsigned grid, bool count, cudaStream_t st);
