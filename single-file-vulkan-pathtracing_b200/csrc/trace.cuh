// trace.cuh — launch interface of the traversal kernel (trace.cu).
#pragma once
#include "common.cuh"
#include "shade.cuh"

constexpr int kTraceBlock = 1024;      // one persistent CTA of 32 warps per SM (four CTAs of 8 warps measured -2.4 %)
constexpr int kTraceSmemStack = 8;     // per-lane stack entries held in shared memory
constexpr int kTraceMaxSmem = 227 * 1024;
// Fused path kernel (k_trace<..., FUSED>): every warp owns kFusedSlots path slots of 64 bytes in shared memory — ray,
// 1/d, octant | depth, throughput, seed, path id, and the closest hit once the ray is done — plus the lists of the
// slots whose ray is finished ("done": waiting to be shaded) and of those that hold a fresh ray ("ready")
constexpr int kFusedSlots = 64;
static_assert((kFusedSlots & (kFusedSlots - 1)) == 0 && kFusedSlots <= 128, "slot ids are masked with kFusedSlots - 1 and kept in bytes");
constexpr int kFusedArgsBytes = 256;                              // copy of FusedArgs
constexpr int kFusedWarpBytes = 64 + 64 + 64 + kFusedSlots * 64;  // done list, ready list, counters | slots
// mbarrier + octant permutation table + (wavefront instance: per-warp ray pools | fused instance: arguments + per-warp path slots)
__host__ __device__ constexpr int trace_smem_fixed(int block, bool fused = false) {
    return 16 + 2048 + (fused ? kFusedArgsBytes + (block / 32) * kFusedWarpBytes : block * 48);
}
constexpr int kTraceLocalStack = 40;   // overflow entries per lane in local memory
// every node step pushes at most one sibling group and one parked triangle group
constexpr int kTraceMaxDepth = (kTraceSmemStack + kTraceLocalStack) / 2 - 1;

// device counters (bpt_stats): rays always; the rest only from the instrumented kernel
enum { BPT_STAT_RAYS = 0, BPT_STAT_NODES, BPT_STAT_TRIS, BPT_STAT_WARP_ITERS, BPT_STAT_WARP_NODE_STEPS,
       BPT_STAT_WARP_TRI_STEPS, BPT_STAT_LANE_ITERS, BPT_STAT_COUNT };

// What the fused instance needs on top of the traversal arguments: it generates the primary rays of a sample pass
// (raygen.rgen:47-57), traces them, shades every hit (shade_one.cuh) and traces the bounce ray, until the pass has no
// path left; nothing but the per-path colour ever leaves the SM.
struct FusedArgs {
    FrameParams p;
    SceneView s;
    float4* path_color;        // per path id of the pass (shade.cuh)
    uint32_t* path_ctr;        // next path of the pass, zeroed before the launch
    const int32_t* frame_dev;  // null, or the device int that overrides p.frame (graph replay)
    uint32_t s0;               // first sample of the pass
    uint32_t npaths;           // tile pixels * samples of the pass
    uint32_t npix;             // tile pixels
    uint32_t path_base;        // path id of the pass's first path
};
static_assert(sizeof(FusedArgs) <= kFusedArgsBytes, "FusedArgs must fit its shared-memory copy");

struct TraceArgs {
    const float4* rays;            // 2 per ray: {o, tmin} {d, tmax}
    uint4* hits;                   // {t, -, -, prim or BPT_MISS}
    const uint32_t* count_ptr;     // number of rays (device)
    uint32_t* fetch_ctr;           // global ray fetch counter (zeroed before launch)
    const Node8* recs;             // the record array: BVH8 nodes and triangle / instance records (common.cuh)
    uint32_t staged_recs;          // STAGED instance: number of records (all of them) to copy into shared memory
    uint32_t root;                 // record of the root node (0; the instance-level root in two-level scenes — every
                                   // record from it on belongs to the instance level)
    uint32_t num_mesh_tris;        // TWO_LEVEL: primitive id = instance * num_mesh_tris + triangle
    float gbias[2][3], gstep[2][3];  // scene grids of the node origins ([0] mesh level, [1] instance level):
                                   // origin = fmaf(float(2^23 + c), gstep, gbias), gbias = grid_lo - 2^23 * gstep
    int refill_below;              // refill a warp's idle lanes when fewer than this many are live
    int steps_per_refill;          // traversal iterations between two refill votes
    int staged_tris_per_step;      // STAGED instance: triangle tests per lane per iteration (shared-memory records)
    uint32_t magic;                // 0x47000000 (float 32768): byte->float permute constant, see trace.cu byte_f
    unsigned long long* stat;      // BPT_STAT_* counters (may be null when not counting: only [RAYS] is touched)
    int count_rays;                // 1: add the launch's ray count to stat[RAYS] (0: shadow-ray launches, counted by k_nee)
    FusedArgs f;                   // fused instance only
};

size_t trace_smem_bytes(uint32_t staged_recs, int block = kTraceBlock, bool fused = false);
cudaError_t trace_configure();
// fused: the path kernel (a.f filled in; a.rays / a.hits / a.count_ptr / a.fetch_ctr unused)
void trace_launch(const TraceArgs& a, unsigned num_sms, bool staged, bool two_level, bool count, bool fused, cudaStream_t st);
