/*
 * bpt.h — C ABI of the B200-native wavefront path tracer ("bpt").
 *
 * This is the drop-in boundary for the ONE hot path of
 * yknishidate/single-file-vulkan-pathtracing: "build an acceleration structure over the
 * uploaded triangle buffers, then trace W x H x spp paths and accumulate radiance".
 * Every entry point names the reference call site it replaces (paths relative to the
 * reference checkout, e.g. main.cpp:659).
 *
 * Conventions
 *   - plain C types only; no CUDA, torch or C++ types cross this boundary
 *     (streams / device pointers travel as void*).
 *   - every call returns 0 on success, a negative BPT_E_* code on failure; the message is
 *     available from bpt_last_error(ctx). No C++ exception crosses the ABI
 *     (the reference throws std::runtime_error, main.cpp:35,63,118,151,221,594,612,681).
 *   - a context is bound to one GPU and one CUDA stream and is externally synchronised
 *     (the reference is single-threaded with one queue, main.cpp:171,683).
 *   - calls are stream-ordered and asynchronous unless they return host data.
 *   - there is NO CPU fallback: without a CUDA device bpt_create fails.
 */
#ifndef BPT_H_
#define BPT_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPT_ABI_VERSION 2

/* error codes */
#define BPT_OK            0
#define BPT_E_INVALID    -1  /* bad argument / call order                           */
#define BPT_E_CUDA       -2  /* a CUDA runtime call failed (message has the detail) */
#define BPT_E_NCCL       -3  /* NCCL missing or a NCCL call failed                  */
#define BPT_E_NOMEM      -4
#define BPT_E_STATE      -5  /* e.g. trace before build                             */

/* accumulate modes (raygen.rgen:86-90) */
#define BPT_ACCUM_FLOAT4  0  /* running mean kept in a float4 image (north-star default)          */
#define BPT_ACCUM_RGBA8   1  /* running mean re-quantised to unorm8 every frame, exactly what the
                                reference's B8G8R8A8Unorm storage image does (main.cpp:481-484)   */

/* sampler modes */
#define BPT_SAMPLER_UNIFORM 0 /* raygen.rgen:23-30,79 — uniform hemisphere, pdf 1/2pi (PARITY)    */
#define BPT_SAMPLER_COSINE  1 /* opt-in, NOT same-seed comparable with the reference              */
/* Non-parity estimator switches (bpt_params.rr_start_depth, .nee; both 0 = the reference's estimator). They estimate
 * the same integral as the reference's loop (raygen.rgen:62-84, truncated at max_depth segments) with less variance
 * or less work, consume extra random numbers, and are therefore validated by convergence, never by same-seed equality:
 *   rr_start_depth = k > 0 : Russian roulette after the hit of segment d >= k-1 (0-based): the path survives with
 *       probability q = min(1, max(w.r, w.g, w.b)) of its updated weight (one more rand(seed)) and w /= q.
 *   nee = 1 : next-event estimation. At every hit whose next segment would be traced, one point on the emissive
 *       triangles (Ke != 0; chosen with probability proportional to area, uniform on the triangle: three rand(seed)
 *       before the two of the bounce) is connected by a shadow ray; Ke of a triangle hit by a BOUNCE ray is then not
 *       added again (camera rays still add it). Instanced scenes: every instance of an emissive triangle is a light. */

typedef struct bpt_context bpt_context; /* opaque; owns all device memory */

/*
 * Launch parameters of one frame. The reference hard-codes all of these
 * (main.cpp:16-17, raygen.rgen:43,55-56,62,71,73, miss.rmiss:10); bpt_params_default()
 * returns exactly those constants.
 */
typedef struct bpt_params {
    uint32_t width;          /* full image width  (gl_LaunchSizeEXT.x, main.cpp:659)   */
    uint32_t height;         /* full image height (gl_LaunchSizeEXT.y)                 */
    uint32_t spp_per_frame;  /* maxSamples, raygen.rgen:43 (32)                        */
    uint32_t max_depth;      /* segment bound, raygen.rgen:62 (8)                      */
    int32_t  frame;          /* push constant `frame`, main.cpp:658                    */
    uint32_t tile_y0;        /* first image row this context renders (contiguous tile) */
    uint32_t tile_rows;      /* number of rows; 0 = all rows from tile_y0              */
    float    cam_origin[3];  /* raygen.rgen:55 (0,-1,5)                                */
    float    cam_target[3];  /* target = (d.x + t[0], d.y + t[1], t[2]); raygen.rgen:56 (0,-1,2) */
    float    sky[3];         /* miss.rmiss:10 (0.7,0.6,0.5)                            */
    float    tmin;           /* raygen.rgen:71 (0.001)                                 */
    float    tmax;           /* raygen.rgen:73 (10000)                                 */
    uint32_t accum_mode;     /* BPT_ACCUM_*                                            */
    uint32_t sampler;        /* BPT_SAMPLER_*                                          */
    /* Interleaved multi-GPU tiling (tile_block != 0; tile_y0/tile_rows must then be 0): the
     * image is cut into blocks of tile_block rows, dealt round-robin to tile_nranks contexts;
     * this context renders the blocks b with b % tile_nranks == tile_rank, so sky rows and
     * geometry rows spread evenly over the GPUs. height % (tile_block*tile_nranks) == 0.
     * Its rows are stored contiguously ("rank-major") at row tile_rank*height/tile_nranks of
     * the image buffer, which is what lets one in-place all-gather assemble the image;
     * bpt_read_image* return ordinary row-major images in every mode. */
    uint32_t tile_block;
    uint32_t tile_nranks;
    uint32_t tile_rank;
    uint32_t rr_start_depth; /* 0 = off (reference); see "Non-parity estimator switches"   */
    uint32_t nee;            /* 0 = off (reference)                                         */
} bpt_params;

/* Counters of the last bpt_trace / since bpt_reset_stats. */
typedef struct bpt_stats {
    uint64_t rays_traced;     /* sum over bounces of live-queue length (device counted)     */
    uint64_t paths;           /* tile pixels * spp traced                                   */
    uint64_t trace_launches;  /* launches of the traversal kernel                           */
    uint64_t kernel_launches; /* all kernel launches issued by bpt_trace                    */
    double   trace_kernel_ms; /* device time inside the traversal kernel (CUDA events);
                                 only filled when profiling is on (bpt_set_option)          */
    double   frame_ms;        /* device time of the whole bpt_trace call(s) (CUDA events)   */
    double   build_ms;        /* device time of the last bpt_build_accel                    */
    uint64_t nodes_visited;   /* instrumented builds only (BPT_OPT_COUNT_TRAVERSAL)         */
    uint64_t tris_tested;     /*   "                                                        */
    /* SIMD efficiency of the traversal loop (instrumented builds only): one loop iteration of a
     * warp does at most one node step and one triangle test per lane.
     *   lanes per node step     = nodes_visited / warp_node_steps
     *   lanes per triangle step = tris_tested   / warp_tri_steps
     *   live lanes per iteration = lane_iterations / warp_iterations                           */
    uint64_t warp_iterations;
    uint64_t warp_node_steps;
    uint64_t warp_tri_steps;
    uint64_t lane_iterations;
} bpt_stats;

/* Sizes of the acceleration structure, for layout/roofline documentation and tests. */
typedef struct bpt_accel_info {
    uint32_t num_tris;        /* triangles in the bottom-level structure                    */
    uint32_t num_instances;
    uint32_t num_nodes8;      /* BVH8 nodes (64 B each)                                     */
    uint32_t num_binary_nodes;/* LBVH internal nodes (N-1)                                  */
    uint32_t top_nodes_smem;  /* records (nodes + triangles) staged into shared memory by TMA:
                                 all of them for a scene small enough, else 0               */
    uint32_t max_depth8;      /* depth of the BVH8                                          */
    uint64_t bytes_nodes;     /* num_nodes8 * 64                                            */
    uint64_t bytes_tris;      /* num_tris * 64 (Woop rows + primitive id)                   */
    uint32_t num_tlas_nodes8; /* two-level builds: nodes in the instance BVH8               */
    uint32_t num_records;     /* nodes + triangle records of the (mesh-level) record array  */
} bpt_accel_info;

/* options for bpt_set_option */
#define BPT_OPT_PROFILE          1 /* 1: bracket every traversal launch with CUDA events     */
#define BPT_OPT_COUNT_TRAVERSAL  2 /* 1: use the instrumented traversal kernel (nodes/tris)  */
#define BPT_OPT_SMEM_TOP_NODES   3 /* stage the whole BVH in shared memory (TMA) when it has at most
                                      this many nodes and fits (0 = never stage)             */
#define BPT_OPT_STREAMS          4 /* sample lanes per pass (1..4, default 1): the samples of a pass are split into this many
                                      independent wavefronts on their own CUDA streams, so that the tail of one lane's persistent
                                      traversal launch can be filled by the other lanes' kernels; results do not depend on it.
                                      Measured on B200: no gain, every kernel fills the whole machine */
#define BPT_OPT_USE_GRAPH        6 /* 1: capture a frame's launch list as a CUDA graph and replay it while only the
                                      frame index changes (pays off for launch-bound, small frames)             */
#define BPT_OPT_FUSED_PATHS      13 /* 1: bpt_trace runs ONE path kernel per sample pass — primary rays, traversal, closest-hit /
                                     miss and the bounce all inside the SMs, path state in shared memory — instead of the
                                     per-bounce wavefront (generate, then traverse + shade per bounce through ray / hit queues
                                     in HBM). Same paths, same arithmetic: the image is bit-identical. Default 0: on B200 it
                                     is 10-20 % faster for frames of up to ~2 M paths (1024 x 1024 at 1 spp) and 6-10 % slower
                                     for the big ones. The wavefront always runs the estimators the reference does not have
                                     (nee, rr_start_depth), the sample lanes (BPT_OPT_STREAMS > 1) and staged scenes too big to
                                     leave room for the path slots */
#define BPT_OPT_PASS_PATHS       10 /* target paths per sample pass: a pass carries min(spp, this / tile pixels)
                                      samples of every tile pixel (default 2^27); results do not depend on it */
#define BPT_OPT_BVH_OPTIMAL_COLLAPSE 7 /* 1 (default): SAH-optimal binary -> 8-wide collapse (Ylitie et al. 2017, dynamic
                                        programme); 0: greedy, largest surface area first                          */
#define BPT_OPT_BVH_SAH_SUBTREE  12 /* quality stage of the build (stands in for ePreferFastTrace, main.cpp:419): every subtree of
                                       the binary LBVH with at most this many triangles (3..32; default 32; 0 = plain LBVH) is
                                       rebuilt with the full-sweep surface-area heuristic before the collapse to 8-wide nodes */
#define BPT_OPT_TRACE_REFILL_BELOW 8     /* traversal: refill a warp when fewer lanes than this are live */
#define BPT_OPT_TRACE_STEPS_PER_REFILL 9 /* traversal: loop iterations between two refill votes           */
#define BPT_OPT_TRACE_STAGED_TRIS_PER_STEP 11 /* traversal of a shared-memory-staged scene: triangle tests per lane
                                                per loop iteration (big scenes always run one)            */

/* ---- lifecycle: replaces Context ctor/dtor (main.cpp:74-267) ------------------------- */
int  bpt_abi_version(void);
void bpt_params_default(bpt_params* p);
/* device: CUDA ordinal. stream: a cudaStream_t to enqueue on, or NULL to create one. */
int  bpt_create(int device, void* stream, bpt_context** out);
void bpt_destroy(bpt_context* ctx);
const char* bpt_last_error(const bpt_context* ctx); /* ctx may be NULL: last create error */
int  bpt_set_option(bpt_context* ctx, int option, int64_t value);

/* ---- scene upload: replaces the three Buffer(...) constructions (main.cpp:492-494) ----
 * verts   : xyz float triples, stride 12 B   (struct Vertex, main.cpp:19-21)
 * indices : uint32, 3 per triangle           (main.cpp:45)
 * faces   : Kd.rgb, Ke.rgb per triangle, stride 24 B (struct Face, main.cpp:23-26)
 * Host arrays are copied during the call (like the memcpy at main.cpp:321-325). */
int bpt_upload_mesh(bpt_context* ctx,
                    const float* verts, uint32_t nverts,
                    const uint32_t* indices, uint32_t nindices,
                    const float* faces, uint32_t nfaces);
/* Same, but the arrays are already device pointers on ctx's GPU (zero-copy adoption is not
 * implied: they are copied device-to-device). */
int bpt_upload_mesh_device(bpt_context* ctx,
                           const void* d_verts, uint32_t nverts,
                           const void* d_indices, uint32_t nindices,
                           const void* d_faces, uint32_t nfaces);
/* The scene front-end on the device: what the body of the reference's loadFromFile does on the host (main.cpp:37-57),
 * from the arrays tinyobj::LoadObj returns (main.cpp:34): every face corner becomes its own vertex with Y negated
 * (:41-44), the index buffer becomes 0,1,2,... (:45), every face takes {Kd, Ke} of its material (:47-56).
 *   positions       : attrib.vertices, xyz triples (npositions of them)
 *   corner_vertex   : index_t::vertex_index of every face corner, all shapes in order (ncorners, a multiple of 3)
 *   face_material   : mesh.material_ids of every face (ncorners / 3 of them)
 *   materials_kd_ke : diffuse.rgb, emission.rgb per material (nmaterials of them)
 * A corner that names no position or a face without a valid material is an error (the reference throws on the latter,
 * main.cpp:49-51). Equivalent to bpt_upload_mesh of the arrays loadFromFile produces. */
int bpt_upload_obj_arrays(bpt_context* ctx, const float* positions, uint32_t npositions, const int32_t* corner_vertex,
                          uint32_t ncorners, const int32_t* face_material, const float* materials_kd_ke, uint32_t nmaterials);
/* Instance transforms: n row-major 3x4 float matrices (VkTransformMatrixKHR, main.cpp:515-520), object -> world.
 * Call after bpt_upload_mesh (which resets the scene to one identity instance) and before bpt_build_accel: the
 * build then adds an instance-level BVH8 over the instances' world boxes (the reference's TLAS, main.cpp:538) and
 * the traversal kernel runs two-level. Hit primitive ids are instance * ntris + primitive. Shading transforms the
 * triangle's vertices by the instance matrix and applies the reference's formulas to them (the reference shader
 * itself ignores instance transforms, which is only correct for its single identity instance).
 * Fails on a singular or non-finite matrix. */
int bpt_set_instances(bpt_context* ctx, const float* xforms3x4, uint32_t n);

/* Synthetic triangle soup generated on the device (no reference equivalent —
 * bench/test input only). Equivalent to bpt_upload_mesh of the oracle's orc_soup output. */
int bpt_upload_soup(bpt_context* ctx, uint32_t ntris, uint32_t seed);

/* ---- build: replaces Accel(...) -> buildAccelerationStructuresKHR (main.cpp:416-450,
 *      called for the BLAS at :512 and the TLAS at :538). The mesh level is rebuilt only after a new mesh was
 *      uploaded; bpt_set_instances + bpt_build_accel rebuilds the instance level alone (moving instances).
 *      Like the reference's build (oneTimeSubmit + queue.waitIdle, main.cpp:224-238) the call returns when the
 *      structure is complete: the host reads the scene bounds and the per-level node counts back. */
int bpt_build_accel(bpt_context* ctx);
int bpt_accel_info_get(bpt_context* ctx, bpt_accel_info* out);

/* ---- trace: replaces pushConstants(frame) + traceRaysKHR(..., W, H, 1)
 *      (main.cpp:658-659) and everything raygen/closesthit/miss do per frame. Enqueues the
 *      whole wavefront launch loop on the stream; no host synchronisation. */
int bpt_trace(bpt_context* ctx, const bpt_params* p);

/* host-visible completion: replaces queue.waitIdle() (main.cpp:237,683) */
int bpt_sync(bpt_context* ctx);

/* ---- output: replaces the storage image bound at binding 1 (main.cpp:481-484,637) ---- */
/* rgba: width*height*4 floats, row-major, full image (rows outside the tile are whatever
 * the image buffer holds: zero, or the gathered tiles after bpt_allgather_image). Syncs. */
int bpt_read_image(bpt_context* ctx, float* rgba, size_t nfloats);
/* BGRA8 view as the reference's B8G8R8A8Unorm image would hold it (main.cpp:483). Syncs. */
int bpt_read_image_bgra8(bpt_context* ctx, uint8_t* bgra, size_t nbytes);
/* Present path without a stall (main.cpp:661-683 copies the image to the swapchain and then waits for the queue;
 * here the copy of frame f overlaps the trace of frame f+1): enqueues a device->host copy of the image as it is after
 * everything enqueued so far, on a copy stream of the context. The next bpt_trace may be enqueued at once; the kernel
 * that writes the image again (and bpt_allgather_image) waits for the copy on the device. `rgba` / `bgra` must stay valid
 * (and should be pinned) until bpt_read_wait returns. One copy may be pending per context. */
int bpt_read_image_async(bpt_context* ctx, float* rgba, size_t nfloats);
int bpt_read_image_bgra8_async(bpt_context* ctx, uint8_t* bgra, size_t nbytes);
int bpt_read_wait(bpt_context* ctx); /* host-visible completion of the pending copy (no-op if none) */
/* device pointer of the float4 image (valid until the next resize); for interop/present. */
int bpt_image_device_ptr(bpt_context* ctx, void** dptr, size_t* nbytes);
/* zero the image and restart accumulation (a fresh outputImage). */
int bpt_clear_image(bpt_context* ctx);

int bpt_get_stats(bpt_context* ctx, bpt_stats* out); /* syncs the stream */
int bpt_reset_stats(bpt_context* ctx);

/* ---- stage-level entry points (what traceRayEXT alone does, raygen.rgen:63-75).
 * Used by the parity tests and by callers that bring their own rays.
 * rays : n records of 8 floats {ox,oy,oz,tmin, dx,dy,dz,tmax} (host memory)
 * hits : n records of 4 words  {t, u, v, prim}; prim = 0xffffffff on miss; for
 *        multi-instance scenes prim = instance * ntris + primitive. */
int bpt_trace_rays(bpt_context* ctx, const float* rays, uint32_t n, void* hits);
/* One closest-hit / miss + path-update step for n paths, alone (closesthit.rchit:50-65, miss.rmiss:8-12,
 * raygen.rgen:76-83): the stage-level view of the shade kernel, for parity tests. Host arrays:
 *   in : rays n*8 (as above), hits n*4 words (only t and prim are read: u, v are re-derived from the original
 *        vertices as the frame loop does), weight n*3, seed n
 *   out: contrib n*3 (= weight * emission or weight * sky: what raygen.rgen:76 adds to `color`), new_rays n*8,
 *        new_weight n*3, new_seed n, alive n bytes (0: the path ended on a miss; its other outputs are zero)
 * p supplies sky, tmin, tmax and the sampler; the next segment is always sampled (as raygen.rgen:78-80 does). */
int bpt_shade_step(bpt_context* ctx, const bpt_params* p, uint32_t n, const float* rays, const void* hits,
                   const float* weight, const uint32_t* seed, float* contrib, float* new_rays, float* new_weight,
                   uint32_t* new_seed, uint8_t* alive);
/* primary rays + seeds of sample `sample_in_frame` for the tile in p
 * (raygen.rgen:47-57). rays: tile_pixels*8 floats, seeds: tile_pixels uint32 (post-jitter
 * RNG state). */
int bpt_generate_rays(bpt_context* ctx, const bpt_params* p, uint32_t sample_in_frame,
                      float* rays, uint32_t* seeds);

/* ---- BVH introspection for the build-invariant tests (copies device -> host) --------- */
/* The mesh-level record array: records = num_records * 64 bytes (BVH8 nodes and triangle records interleaved, record
 * 0 = root; layouts in csrc/common.cuh), rec_prim = num_records uint32 (primitive id of a triangle record, 0xffffffff
 * for a node record), grid = 6 floats (bias xyz and step xyz of the grid the node origins are quantised on:
 * origin = fmaf(float(2^23 + c), step, bias)).
 * Any pointer may be NULL. */
int bpt_download_accel(bpt_context* ctx, void* records, uint32_t* rec_prim, float* grid);
/* the uploaded (or device-generated) mesh, as bpt_upload_mesh would have received it. */
int bpt_download_mesh(bpt_context* ctx, float* verts, uint32_t* indices, float* faces);
/* sorted 64-bit (morton<<32 | prim) keys of the last build. */
int bpt_download_morton(bpt_context* ctx, uint64_t* keys, uint32_t n);
/* primitive of every leaf of the binary hierarchy, in leaf order: the sorted order of the keys, permuted inside the
 * subtrees the SAH stage rebuilt (BPT_OPT_BVH_SAH_SUBTREE). */
int bpt_download_leaf_order(bpt_context* ctx, uint32_t* prims, uint32_t n);
/* binary LBVH: parent-less arrays left[n-1], right[n-1] (child >= n-1 means leaf
 * child-(n-1)), and aabbs[(2n-1)*6] (internal nodes first, then leaves). */
int bpt_download_lbvh(bpt_context* ctx, uint32_t* left, uint32_t* right, float* aabbs);

/* ---- multi-GPU (no reference equivalent: it uses enumeratePhysicalDevices().front(),
 *      main.cpp:105). One context per process/GPU; NCCL is dlopen'ed at first use. ------ */
#define BPT_NCCL_UNIQUE_ID_BYTES 128
int bpt_nccl_unique_id(uint8_t id[BPT_NCCL_UNIQUE_ID_BYTES]);
int bpt_nccl_init(bpt_context* ctx, const uint8_t id[BPT_NCCL_UNIQUE_ID_BYTES],
                  int rank, int nranks);
/* In-place all-gather of the row tiles into every rank's full float4 image. Tiles must be
 * either the equal contiguous split bpt_tile_rows() returns (height % nranks == 0) or the
 * interleaved split of bpt_params.tile_block with tile_nranks == nranks, tile_rank == rank
 * (the tiling of the last bpt_trace decides). */
int bpt_allgather_image(bpt_context* ctx, uint32_t width, uint32_t height);
/* contiguous row split used by every caller: rank r gets rows [y0, y0+rows), height / nranks each; the last rank
 * also takes the height % nranks remaining rows (bpt_allgather_image needs an even split and says so). */
void bpt_tile_rows(uint32_t height, int rank, int nranks, uint32_t* y0, uint32_t* rows);

#ifdef __cplusplus
}
#endif
#endif /* BPT_H_ */
