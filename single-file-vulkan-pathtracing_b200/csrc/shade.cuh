// shade.cuh — launch interface of the wavefront stages around the traversal kernel (shade.cu).
#pragma once
#include "common.cuh"

// Device-resident parameters of the running frame (the reference's push constant + the constants
// it hard-codes in the shaders).
struct FrameParams {
    uint32_t width, height, spp_per_frame, max_depth;
    int32_t frame;
    uint32_t tile_y0, tile_rows;
    float cam_origin[3], cam_target[3], sky[3];
    float tmin, tmax;
    uint32_t accum_mode, sampler;
    uint32_t tile_block, tile_nranks, tile_rank;
    uint32_t rr_start_depth, nee;
};

// Row bookkeeping of a tile (bpt_params.tile_*): l is the tile-local row.
__host__ __device__ __forceinline__ uint32_t tile_local_rows(const FrameParams& p) {
    if (p.tile_block) return p.height / p.tile_nranks;
    return p.tile_rows ? p.tile_rows : p.height - p.tile_y0;
}
__host__ __device__ __forceinline__ uint32_t tile_global_row(const FrameParams& p, uint32_t l) {  // image row (seeds, camera)
    if (p.tile_block) return ((l / p.tile_block) * p.tile_nranks + p.tile_rank) * p.tile_block + l % p.tile_block;
    return p.tile_y0 + l;
}
__host__ __device__ __forceinline__ uint32_t tile_store_row(const FrameParams& p, uint32_t l) {   // row of the image buffer
    if (p.tile_block) return p.tile_rank * (p.height / p.tile_nranks) + l;
    return p.tile_y0 + l;
}

struct SceneView {
    const float4* srec;       // 4 x float4 per primitive: v0 v1 v2 Kd Ke - (built from the three arrays of
                              // main.cpp:492-494 by launch_shade_records)
    const float* xforms;      // 3x4 row-major per instance, or null (single identity instance)
    uint32_t ntris;
};
void launch_shade_records(const float* verts, const uint32_t* idx, const float* faces, uint32_t ntris, float4* out, cudaStream_t st);

// Next-event estimation (bpt_params.nee, include/bpt.h; not the reference's estimator): the emissive triangles in primitive
// order with the cumulative distribution of their areas, and per path id the solid-angle pdf its current segment's direction
// was sampled with (the balance heuristic needs it when a bounce ray finds an emitter).
struct NeeView {
    const uint32_t* light_prims;
    const float* light_cdf;   // float(cumulative area / total area), last entry 1
    uint32_t nlights;
    float light_area;         // total area of the emissive triangles
    float* pdf_prev;          // null when next-event estimation is off
};

// Per-pass device counters: counts[d] = length of bounce d's queue, fetch[d] = ray fetch counter of bounce d's traversal
// launch, fetch[kCounterStride + d] = tile counter of bounce d's shade launch, fetch[2 * kCounterStride + d] = ray fetch
// counter of bounce d's shadow-ray launch (next-event estimation); k_generate resets all four.
constexpr uint32_t kCounterStride = 65;  // kMaxDepth + 1 (api.cu)

// One wavefront queue (SoA): ray 2 x float4, state float4 {w.rgb, seed bits}, pixel u32 (= path id of the pass).
struct PathQueue {
    float4* rays;
    float4* state;
    uint32_t* pixel;
};

// one lane of a pass = samples s0..s0+ns-1 of every tile pixel; path id = path_base + slot * tile_pixels + tile-local
// pixel (path_base = the lane's first sample slot of the pass * tile_pixels; q, counts, fetch are the lane's own)
// frame_dev (may be null): device int that overrides p.frame (graph replay)
void launch_generate(const FrameParams& p, const int32_t* frame_dev, uint32_t s0, uint32_t ns, uint32_t path_base, PathQueue q,
                     uint32_t* counts, uint32_t* fetch, uint32_t ncounters, cudaStream_t st);
void launch_set_i32(int32_t* dst, int32_t v, cudaStream_t st);
// folds the per-sample colours of a finished pass into the frame sum (sample order) and clears them
void launch_gather_pass(uint32_t npix, uint32_t ns, float4* path_color, float4* frame_sum, cudaStream_t st);
// depth: index of the bounce being shaded; counts[depth] paths in `in`, survivors appended to `out`
// and counted in counts[depth+1]. path_color: per-path radiance of the pass (indexed by path id).
void launch_shade(const FrameParams& p, const SceneView& s, const NeeView& nv, uint32_t depth, PathQueue in, const uint4* hits,
                  PathQueue out, uint32_t* counts, uint32_t* fetch, float4* path_color, uint32_t max_paths, unsigned num_sms,
                  cudaStream_t st);
// Next-event estimation, run between the traversal and the shade of bounce `depth`: for every path of `in` that hit
// something, one shadow ray towards an area-sampled point of the emissive triangles (shadow_rays 2 x float4 per path;
// tmax < 0 marks "no connection") and the radiance it carries if nothing is in the way (shadow_contrib); consumes three
// rand(seed) per hit (in.state is updated in place) and adds the number of shadow rays to *ray_stat. After the shadow
// rays were traced, launch_nee_resolve adds the contributions of the unoccluded ones to the paths' colours.
void launch_nee(const FrameParams& p, const SceneView& s, const NeeView& nv, uint32_t depth, PathQueue in, const uint4* hits,
                const uint32_t* counts, float4* shadow_rays, float4* shadow_contrib, unsigned long long* ray_stat,
                uint32_t max_paths, unsigned num_sms, cudaStream_t st);
void launch_nee_resolve(uint32_t depth, const uint32_t* counts, const uint4* shadow_hits, const float4* shadow_contrib,
                        const uint32_t* pixel, float4* path_color, uint32_t max_paths, unsigned num_sms, cudaStream_t st);
// re-derives (u,v) of every hit from the original vertices (what k_shade does internally)
void launch_refine_hits(const SceneView& s, const float4* rays, uint4* hits, uint32_t n, cudaStream_t st);
void launch_accumulate(const FrameParams& p, const int32_t* frame_dev, float4* frame_sum, float4* image, cudaStream_t st);
// the reference's loadFromFile body on the device (main.cpp:37-57); *bad counts corners / faces with an invalid reference
void launch_obj_arrays(const float* positions, uint32_t npositions, const int32_t* corner_vertex, uint32_t ncorners,
                       const int32_t* face_material, const float* materials, uint32_t nmaterials, float* verts, uint32_t* idx,
                       float* faces, uint32_t* bad, cudaStream_t st);
// out[0] = number of indices >= nverts, out[1] = position of the first one (out zeroed by the caller)
void launch_check_indices(const uint32_t* idx, uint32_t n, uint32_t nverts, uint32_t* out, cudaStream_t st);
void launch_soup(uint32_t ntris, uint32_t seed, float scale, float* verts, uint32_t* idx, float* faces, cudaStream_t st);
void launch_image_to_bgra8(const float4* image, uint8_t* bgra, size_t npix, cudaStream_t st);
// rank-major (interleaved tiling) image buffer -> row-major image
void launch_deinterleave(const float4* rank_major, float4* row_major, uint32_t width, uint32_t height, uint32_t block,
                         uint32_t nranks, cudaStream_t st);
