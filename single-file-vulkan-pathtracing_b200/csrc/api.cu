// api.cu — the extern "C" boundary of libbpt (include/bpt.h) and the host-side launch loop.
//
// The launch loop in bpt_trace is what replaces the reference's single
// vkCmdTraceRaysKHR(raygen, miss, hit, {}, W, H, 1) (main.cpp:659): per sample pass one
// `generate`, then per bounce one persistent `trace` + one `shade` (with ballot compaction into
// the next queue), and one `accumulate` per frame — all enqueued on one CUDA stream, no host
// synchronisation inside.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "build.cuh"
#include "nccl_dl.h"
#include "shade.cuh"
#include "trace.cuh"

namespace {
constexpr uint32_t kMaxDepth = 64;
static_assert(kCounterStride == kMaxDepth + 1, "counter arrays are kMaxDepth + 1 long");
constexpr int kMaxLanes = 4;                              // BPT_OPT_STREAMS
constexpr uint32_t kLaneCounters = 4 * (kMaxDepth + 1);   // counts[], fetch[], shade tile counters[], shadow fetch[] of one lane
std::string g_create_error;
}  // namespace

struct bpt_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = BPT_NUM_SMS_DEFAULT;
    std::string err;

    // scene (copies of the caller's arrays, main.cpp:492-494)
    float* d_verts = nullptr;
    uint32_t* d_idx = nullptr;
    float* d_faces = nullptr;
    float* d_xforms = nullptr;      // null: one identity instance
    float* d_xforms_inv = nullptr;  // inverse 3x4 of every instance (two-level scenes)
    uint32_t nverts = 0, nidx = 0, nfaces = 0, ntris = 0, ninst = 1;

    // acceleration structure
    Bvh8 blas;                      // over the mesh triangles
    Bvh8 tlas;                      // over the instances (two-level scenes only)
    Node8* d_recs_all = nullptr;    // two-level scenes: [mesh records | instance-level records]
    float4* d_srec = nullptr;       // shading records, 64 B per primitive (shade.cuh)
    uint32_t srec_cap = 0;
    bool two_level = false;
    bool built = false, built_nodes_ok = false;
    bool mesh_built = false;  // the mesh-level BVH8, triangle and shading records match the uploaded mesh
    bool staged = false;  // the traversal kernel instance that holds the whole BVH in shared memory is in use
    bool fused_fits = false;  // the fused path kernel's shared memory (path slots + stack + staged records) fits an SM

    // wavefront buffers
    size_t cap_paths = 0;
    PathQueue q[2] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    uint4* hits = nullptr;
    float4* path_color = nullptr;    // per path of the running pass: its sample's colour so far
    float4* frame_sum = nullptr;     // per tile pixel: sum of the finished samples of the frame
    size_t cap_pixels = 0;
    float4* image = nullptr;         // accumulation target; rank-major under interleaved tiling (bpt.h)
    float4* image_linear = nullptr;  // row-major copy produced on demand under interleaved tiling
    uint32_t img_w = 0, img_h = 0;
    uint32_t tile_block = 0, tile_nranks = 1, tile_rank = 0;  // tiling of the last bpt_trace
    // next-event estimation (bpt_params.nee): light table of the uploaded mesh (built on first use) and per-path buffers
    uint32_t* d_light_prims = nullptr;
    float* d_light_cdf = nullptr;
    uint32_t nlights = 0;
    float light_area = 0.f;
    bool lights_built = false;
    float4* shadow_rays = nullptr;           // 2 per path
    float4* shadow_contrib = nullptr;
    uint4* shadow_hits = nullptr;
    float* pdf_prev = nullptr;               // per path id
    size_t cap_nee = 0;
    uint32_t* counters = nullptr;            // per lane: counts[], fetch[], shade tile counters[] (kMaxDepth+1 each; shade.cuh)
    // sample lanes (BPT_OPT_STREAMS): lane 0 runs on `stream`, lane l > 0 on lane_stream[l-1], forked and joined with events
    int num_lanes = 1;
    cudaStream_t lane_stream[kMaxLanes - 1] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes - 1] = {nullptr, nullptr, nullptr};
    // present path: asynchronous read-back on its own stream (bpt_read_image_async)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_img_ready = nullptr, ev_copy_done = nullptr;
    bool copy_pending = false;
    uint8_t* d_bgra = nullptr;               // BGRA8 view of the image, kept with the image (no allocation per read)
    unsigned long long* d_stats = nullptr;   // BPT_STAT_* (trace.cuh)

    // options
    bool profile = false, count = false;
    int64_t opt_stage_max_nodes = 1 << 20;  // BPT_OPT_SMEM_TOP_NODES: 0 disables shared-memory staging
    int refill_below = 30, steps_per_refill = 2, staged_tris_per_step = 2;
    // BPT_OPT_USE_GRAPH: the launch list of a frame captured once as a CUDA graph and replayed while nothing but the
    // frame index changes (the kernels then read the frame index from d_frame)
    bool use_graph = false;
    int32_t* d_frame = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bpt_params graph_key{};
    uint64_t graph_epoch = 0, epoch = 1;  // epoch moves whenever a buffer a captured graph points at may have moved
    uint64_t graph_kernel_launches = 0, graph_trace_launches = 0;
    bool optimal_collapse = true;  // BPT_OPT_BVH_OPTIMAL_COLLAPSE
    uint32_t sah_max_leaves = 32;  // BPT_OPT_BVH_SAH_SUBTREE
    int64_t pass_paths = 1ll << 27;  // BPT_OPT_PASS_PATHS: target number of paths per sample pass
    bool fused = false;  // BPT_OPT_FUSED_PATHS: one path kernel per sample pass instead of the per-bounce wavefront

    // statistics
    bpt_stats stats{};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> frame_events, trace_events;
    std::vector<cudaEvent_t> event_pool;

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
};

int bpt_fail(bpt_context* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}
int bpt_fail_cuda(bpt_context* ctx, cudaError_t e, const char* what, const char* file, int line) {
    return bpt_fail(ctx, BPT_E_CUDA, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
}

namespace {

cudaEvent_t get_event(bpt_context* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Folds the finished event pairs at the head of `list` into `sum_ms` and returns their events to the pool. bpt_trace
// calls it once enough pairs are pending, so that a caller who never asks for statistics (the reference's frame loop
// runs until the window closes, main.cpp:647) does not pile up events; `all`: the stream has been synchronised.
void fold_events(bpt_context* c, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& list, double& sum_ms, bool all) {
    size_t done = 0;
    for (; done < list.size(); ++done) {
        if (!all && cudaEventQuery(list[done].second) != cudaSuccess) {
            (void)cudaGetLastError();  // cudaErrorNotReady is an answer, not a failure of the caller's launches
            break;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, list[done].first, list[done].second);
        sum_ms += ms;
        c->event_pool.push_back(list[done].first);
        c->event_pool.push_back(list[done].second);
    }
    list.erase(list.begin(), list.begin() + (std::ptrdiff_t)done);
}
constexpr size_t kFoldEventsAt = 256;  // pending pairs that trigger a fold in bpt_trace

void free_scene(bpt_context* c) {
    cudaFree(c->d_verts); cudaFree(c->d_idx); cudaFree(c->d_faces); cudaFree(c->d_xforms); cudaFree(c->d_xforms_inv);
    c->d_verts = nullptr; c->d_idx = nullptr; c->d_faces = nullptr; c->d_xforms = nullptr; c->d_xforms_inv = nullptr;
    c->nverts = c->nidx = c->nfaces = c->ntris = 0;
    c->ninst = 1;
    c->built = false;
    c->mesh_built = false;
    c->lights_built = false;
}
void free_paths(bpt_context* c) {
    for (auto& q : c->q) {
        cudaFree(q.rays); cudaFree(q.state); cudaFree(q.pixel);
        q = PathQueue{nullptr, nullptr, nullptr};
    }
    cudaFree(c->hits); cudaFree(c->path_color);
    c->hits = nullptr; c->path_color = nullptr;
    c->cap_paths = 0;
    cudaFree(c->shadow_rays); cudaFree(c->shadow_contrib); cudaFree(c->shadow_hits); cudaFree(c->pdf_prev);
    c->shadow_rays = nullptr; c->shadow_contrib = nullptr; c->shadow_hits = nullptr; c->pdf_prev = nullptr;
    c->cap_nee = 0;
}

// per-path buffers of next-event estimation, as large as the path queues
int ensure_nee(bpt_context* c) {
    if (c->cap_nee >= c->cap_paths) return BPT_OK;
    c->epoch++;
    cudaFree(c->shadow_rays); cudaFree(c->shadow_contrib); cudaFree(c->shadow_hits); cudaFree(c->pdf_prev);
    c->shadow_rays = nullptr; c->shadow_contrib = nullptr; c->shadow_hits = nullptr; c->pdf_prev = nullptr;
    c->cap_nee = 0;
    const size_t n = c->cap_paths;
    BPT_CUDA_TRY(c, cudaMalloc(&c->shadow_rays, n * 2 * sizeof(float4)));
    BPT_CUDA_TRY(c, cudaMalloc(&c->shadow_contrib, n * sizeof(float4)));
    BPT_CUDA_TRY(c, cudaMalloc(&c->shadow_hits, n * sizeof(uint4)));
    BPT_CUDA_TRY(c, cudaMalloc(&c->pdf_prev, n * sizeof(float)));
    c->cap_nee = n;
    return BPT_OK;
}

// Light table of next-event estimation: the emissive triangles (Ke != 0, area > 0) in primitive order and the cumulative
// distribution of their areas. Areas in double from the float vertices, cdf = float(cumulative / total), last entry 1 —
// the definition the CPU checker of the tests uses for the same switch. Built on the host from the uploaded arrays.
int build_light_table(bpt_context* c) {
    if (c->lights_built) return BPT_OK;
    std::vector<float> verts(3 * (size_t)c->nverts), faces(6 * (size_t)c->nfaces);
    std::vector<uint32_t> idx(c->nidx);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    BPT_CUDA_TRY(c, cudaMemcpy(verts.data(), c->d_verts, verts.size() * 4, cudaMemcpyDeviceToHost));
    BPT_CUDA_TRY(c, cudaMemcpy(idx.data(), c->d_idx, idx.size() * 4, cudaMemcpyDeviceToHost));
    BPT_CUDA_TRY(c, cudaMemcpy(faces.data(), c->d_faces, faces.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> xf;
    if (c->d_xforms) {
        xf.resize(12 * (size_t)c->ninst);
        BPT_CUDA_TRY(c, cudaMemcpy(xf.data(), c->d_xforms, xf.size() * 4, cudaMemcpyDeviceToHost));
    }
    std::vector<uint32_t> prims;
    std::vector<double> cum;
    double total = 0.0;
    const uint32_t ninst = c->d_xforms ? c->ninst : 1u;
    for (uint32_t inst = 0; inst < ninst; ++inst) {
        const float* m = c->d_xforms ? &xf[12 * (size_t)inst] : nullptr;
        for (uint32_t t = 0; t < c->ntris; ++t) {
            const float* f = &faces[6 * (size_t)t];
            if (f[3] == 0.0f && f[4] == 0.0f && f[5] == 0.0f) continue;
            float w[3][3];  // world-space vertices as shading sees them: float, x' = m0*x + m1*y + m2*z + m3, left to right
            for (int k = 0; k < 3; ++k) {
                const float* v = &verts[3 * (size_t)idx[3 * (size_t)t + k]];
                for (int r = 0; r < 3; ++r) w[k][r] = m ? m[4 * r] * v[0] + m[4 * r + 1] * v[1] + m[4 * r + 2] * v[2] + m[4 * r + 3] : v[r];
            }
            const double e1[3] = {(double)w[1][0] - w[0][0], (double)w[1][1] - w[0][1], (double)w[1][2] - w[0][2]};
            const double e2[3] = {(double)w[2][0] - w[0][0], (double)w[2][1] - w[0][1], (double)w[2][2] - w[0][2]};
            const double x[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            const double area = 0.5 * std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            if (!(area > 0.0)) continue;
            total += area;
            prims.push_back(inst * c->ntris + t);
            cum.push_back(total);
        }
    }
    std::vector<float> cdf(cum.size());
    for (size_t i = 0; i < cum.size(); ++i) cdf[i] = (float)(cum[i] / total);
    if (!cdf.empty()) cdf.back() = 1.0f;
    cudaFree(c->d_light_prims); cudaFree(c->d_light_cdf);
    c->d_light_prims = nullptr; c->d_light_cdf = nullptr;
    c->nlights = (uint32_t)prims.size();
    c->light_area = (float)total;
    if (c->nlights) {
        BPT_CUDA_TRY(c, cudaMalloc(&c->d_light_prims, prims.size() * 4));
        BPT_CUDA_TRY(c, cudaMalloc(&c->d_light_cdf, cdf.size() * 4));
        BPT_CUDA_TRY(c, cudaMemcpy(c->d_light_prims, prims.data(), prims.size() * 4, cudaMemcpyHostToDevice));
        BPT_CUDA_TRY(c, cudaMemcpy(c->d_light_cdf, cdf.data(), cdf.size() * 4, cudaMemcpyHostToDevice));
    }
    c->lights_built = true;
    c->epoch++;
    return BPT_OK;
}

int ensure_paths(bpt_context* c, size_t n) {
    if (n <= c->cap_paths) return BPT_OK;
    free_paths(c);
    c->epoch++;
    for (auto& q : c->q) {
        BPT_CUDA_TRY(c, cudaMalloc(&q.rays, n * 2 * sizeof(float4)));
        BPT_CUDA_TRY(c, cudaMalloc(&q.state, n * sizeof(float4)));
        BPT_CUDA_TRY(c, cudaMalloc(&q.pixel, n * sizeof(uint32_t)));
    }
    BPT_CUDA_TRY(c, cudaMalloc(&c->hits, n * sizeof(uint4)));
    BPT_CUDA_TRY(c, cudaMalloc(&c->path_color, n * sizeof(float4)));
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->path_color, 0, n * sizeof(float4), c->stream));
    c->cap_paths = n;
    return BPT_OK;
}

int ensure_frame_sum(bpt_context* c, size_t npix) {
    if (npix <= c->cap_pixels) return BPT_OK;
    c->epoch++;
    cudaFree(c->frame_sum);
    c->frame_sum = nullptr;
    c->cap_pixels = 0;
    BPT_CUDA_TRY(c, cudaMalloc(&c->frame_sum, npix * sizeof(float4)));
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->frame_sum, 0, npix * sizeof(float4), c->stream));
    c->cap_pixels = npix;
    return BPT_OK;
}

int ensure_image(bpt_context* c, uint32_t w, uint32_t h) {
    if (c->image && c->img_w == w && c->img_h == h) return BPT_OK;
    c->epoch++;
    if (c->copy_pending) { cudaEventSynchronize(c->ev_copy_done); c->copy_pending = false; }
    cudaFree(c->image); cudaFree(c->image_linear); cudaFree(c->d_bgra);
    c->image = nullptr; c->image_linear = nullptr; c->d_bgra = nullptr;
    BPT_CUDA_TRY(c, cudaMalloc(&c->image, (size_t)w * h * sizeof(float4)));
    BPT_CUDA_TRY(c, cudaMalloc(&c->d_bgra, (size_t)w * h * 4));
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->image, 0, (size_t)w * h * sizeof(float4), c->stream));
    c->img_w = w;
    c->img_h = h;
    return BPT_OK;
}

FrameParams to_frame(const bpt_params* p) {
    FrameParams f;
    static_assert(sizeof(FrameParams) == sizeof(bpt_params), "FrameParams mirrors bpt_params");
    memcpy(&f, p, sizeof(f));
    return f;
}

int check_params(bpt_context* c, const bpt_params* p) {
    if (!p) return bpt_fail(c, BPT_E_INVALID, "params is NULL");
    if (p->width == 0 || p->height == 0) return bpt_fail(c, BPT_E_INVALID, "empty image %ux%u", p->width, p->height);
    if (p->spp_per_frame == 0) return bpt_fail(c, BPT_E_INVALID, "spp_per_frame must be >= 1");
    if (p->max_depth == 0 || p->max_depth > kMaxDepth) return bpt_fail(c, BPT_E_INVALID, "max_depth must be in [1,%u]", kMaxDepth);
    if (p->frame < 0) return bpt_fail(c, BPT_E_INVALID, "frame must be >= 0");
    uint32_t rows = p->tile_rows ? p->tile_rows : p->height - (p->tile_y0 < p->height ? p->tile_y0 : p->height);
    if (p->tile_block) {
        if (p->tile_y0 || p->tile_rows) return bpt_fail(c, BPT_E_INVALID, "interleaved tiling (tile_block != 0) excludes tile_y0/tile_rows");
        if (p->tile_nranks == 0 || p->tile_rank >= p->tile_nranks) return bpt_fail(c, BPT_E_INVALID, "bad tile_rank %u of %u", p->tile_rank, p->tile_nranks);
        if (p->height % ((uint64_t)p->tile_block * p->tile_nranks)) return bpt_fail(c, BPT_E_INVALID, "height %u is not a multiple of tile_block*tile_nranks = %u*%u", p->height, p->tile_block, p->tile_nranks);
        rows = p->height / p->tile_nranks;
    } else if (p->tile_y0 >= p->height || p->tile_y0 + rows > p->height) return bpt_fail(c, BPT_E_INVALID, "tile rows [%u,%u) outside image height %u", p->tile_y0, p->tile_y0 + rows, p->height);
    if ((uint64_t)rows * p->width > 0x7fffffffull) return bpt_fail(c, BPT_E_INVALID, "tile too large");
    if (p->accum_mode > BPT_ACCUM_RGBA8) return bpt_fail(c, BPT_E_INVALID, "bad accum_mode");
    if (p->sampler > BPT_SAMPLER_COSINE) return bpt_fail(c, BPT_E_INVALID, "bad sampler");
    if (p->rr_start_depth > kMaxDepth) return bpt_fail(c, BPT_E_INVALID, "rr_start_depth must be in [0,%u]", kMaxDepth);
    if (p->nee > 1) return bpt_fail(c, BPT_E_INVALID, "bad nee switch");
    return BPT_OK;
}

// Everything that writes the image (or a view of it a pending asynchronous read-back copies from) waits for that copy
// on the device first; the host never blocks.
void wait_pending_copy(bpt_context* c) {
    if (c->copy_pending) cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0);
}

// Row-major view of the image: the buffer itself, or its de-interleaved copy (interleaved tiling).
int row_major_image(bpt_context* c, const float4** out) {
    *out = c->image;
    if (!c->tile_block) return BPT_OK;
    const size_t bytes = (size_t)c->img_w * c->img_h * sizeof(float4);
    if (!c->image_linear) BPT_CUDA_TRY(c, cudaMalloc(&c->image_linear, bytes));
    wait_pending_copy(c);
    launch_deinterleave(c->image, c->image_linear, c->img_w, c->img_h, c->tile_block, c->tile_nranks, c->stream);
    *out = c->image_linear;
    return BPT_OK;
}

TraceArgs make_trace_args(bpt_context* c, const float4* rays, uint4* hits, const uint32_t* count, uint32_t* fetch) {
    TraceArgs a{};
    a.rays = rays; a.hits = hits; a.count_ptr = count; a.fetch_ctr = fetch;
    a.recs = c->two_level ? c->d_recs_all : c->blas.recs;
    a.staged_recs = c->staged ? c->blas.num_recs + (c->two_level ? c->tlas.num_recs : 0u) : 0u;
    a.root = c->two_level ? c->blas.num_recs : 0u;
    a.num_mesh_tris = c->ntris;
    for (int k = 0; k < 3; ++k) {
        a.gbias[0][k] = c->blas.grid_bias[k]; a.gstep[0][k] = c->blas.grid_step[k];
        a.gbias[1][k] = c->two_level ? c->tlas.grid_bias[k] : 0.f; a.gstep[1][k] = c->two_level ? c->tlas.grid_step[k] : 0.f;
    }
    a.refill_below = c->refill_below; a.steps_per_refill = c->steps_per_refill;
    a.staged_tris_per_step = c->staged_tris_per_step;
    a.magic = 0x47000000u;
    a.stat = c->d_stats;
    a.count_rays = 1;
    return a;
}

void launch_trace(bpt_context* c, const TraceArgs& a, cudaStream_t st, bool fused = false) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->profile) {
        e0 = get_event(c); e1 = get_event(c);
        cudaEventRecord(e0, st);
    }
    trace_launch(a, (unsigned)c->num_sms, c->staged, c->two_level, c->count, fused, st);
    if (c->profile) {
        cudaEventRecord(e1, st);
        c->trace_events.emplace_back(e0, e1);
    }
    c->stats.trace_launches++;
    c->stats.kernel_launches++;
}

// The launch list of one frame: per sample pass one generate, per bounce one traversal + one shade, one gather; then
// the running mean. The samples of a pass are dealt to `num_lanes` independent wavefronts ("lanes"), each on its own
// stream with its own slice of the queues and its own counters: a persistent traversal launch ends with a tail in which
// most SMs idle while the last rays finish, and per-launch fixed costs that do not shrink with the tile of a GPU; with a
// second lane in flight the freed SMs pick up the other lane's next kernel at once. A path's colour is collected per
// path id and folded in sample order by the gather, so the image does not depend on the number of lanes.
// frame_dev: null, or the device int the kernels read the frame index from (graph capture).
void enqueue_frame(bpt_context* c, const FrameParams& f, uint32_t npix, uint32_t ns, const int32_t* frame_dev) {
    SceneView sv{c->d_srec, c->d_xforms, c->ntris};
    const bool nee = f.nee && c->nlights > 0;
    const NeeView nv{c->d_light_prims, c->d_light_cdf, c->nlights, c->light_area, f.nee ? c->pdf_prev : nullptr};
    // The fused path kernel (trace.cu, FUSED; opt-in, BPT_OPT_FUSED_PATHS): one launch per sample pass generates, traces
    // and shades every path of the pass inside the SMs; it runs the reference's estimator (and the cosine sampler), the
    // other estimators and the sample lanes stay with the wavefront below. Same paths, same arithmetic, same per-path
    // colours: the image is bit-identical (tests). Measured: 10-20 % faster on frames of up to
    // ~2 M paths (fewer launches and tails), 6-10 % slower on the big ones — the traversal is issue-bound, and the
    // ~1000 instructions of a shade step, which the wavefront's HBM-bound shade kernel hides under its memory traffic,
    // queue up behind it in the same warps.
    const bool fused = c->fused && c->fused_fits && !f.nee && !f.rr_start_depth && c->num_lanes == 1;
    for (uint32_t s0 = 0; fused && s0 < f.spp_per_frame; s0 += ns) {
        const uint32_t n = std::min(ns, f.spp_per_frame - s0);
        uint32_t* path_ctr = c->counters + (kMaxDepth + 1);  // fetch[0] of lane 0
        cudaMemsetAsync(path_ctr, 0, sizeof(uint32_t), c->stream);
        TraceArgs a = make_trace_args(c, nullptr, nullptr, nullptr, nullptr);
        a.f.p = f;
        a.f.s = sv;
        a.f.path_color = c->path_color;
        a.f.path_ctr = path_ctr;
        a.f.frame_dev = frame_dev;
        a.f.s0 = s0;
        a.f.npaths = npix * n;
        a.f.npix = npix;
        a.f.path_base = 0u;
        launch_trace(c, a, c->stream, true);
        launch_gather_pass(npix, n, c->path_color, c->frame_sum, c->stream);
        c->stats.kernel_launches++;
    }
    for (uint32_t s0 = 0; !fused && s0 < f.spp_per_frame; s0 += ns) {
        const uint32_t n = std::min(ns, f.spp_per_frame - s0);
        const uint32_t L = std::min<uint32_t>((uint32_t)c->num_lanes, n);
        if (L > 1) {
            cudaEventRecord(c->ev_fork, c->stream);
            for (uint32_t l = 1; l < L; ++l) cudaStreamWaitEvent(c->lane_stream[l - 1], c->ev_fork, 0);
        }
        struct Lane { cudaStream_t st; PathQueue q[2]; uint4* hits; uint32_t *counts, *fetch; uint32_t paths;
                      float4 *srays, *scontrib; uint4* shits; };
        Lane lane[kMaxLanes];
        uint32_t slot0 = 0;
        for (uint32_t l = 0; l < L; ++l) {
            const uint32_t nl = n / L + (l < n % L ? 1u : 0u);
            const size_t off = (size_t)slot0 * npix;  // first path id of the lane
            Lane& ln = lane[l];
            ln.st = l ? c->lane_stream[l - 1] : c->stream;
            for (int k = 0; k < 2; ++k) ln.q[k] = PathQueue{c->q[k].rays + 2 * off, c->q[k].state + off, c->q[k].pixel + off};
            ln.hits = c->hits + off;
            ln.srays = nee ? c->shadow_rays + 2 * off : nullptr;
            ln.scontrib = nee ? c->shadow_contrib + off : nullptr;
            ln.shits = nee ? c->shadow_hits + off : nullptr;
            ln.counts = c->counters + l * kLaneCounters;
            ln.fetch = ln.counts + (kMaxDepth + 1);
            ln.paths = npix * nl;
            launch_generate(f, frame_dev, s0 + slot0, nl, (uint32_t)off, ln.q[0], ln.counts, ln.fetch, f.max_depth + 1, ln.st);
            c->stats.kernel_launches++;
            slot0 += nl;
        }
        // bounce-major issue order, so that the lanes' kernels alternate in the hardware queues
        for (uint32_t d = 0; d < f.max_depth; ++d)
            for (uint32_t l = 0; l < L; ++l) {
                Lane& ln = lane[l];
                const int cur = (int)(d & 1u);
                launch_trace(c, make_trace_args(c, ln.q[cur].rays, ln.hits, ln.counts + d, ln.fetch + d), ln.st);
                if (nee && d + 1 < f.max_depth) {
                    // next-event estimation: shadow rays of this bounce's hits, traced like any other ray; the unoccluded
                    // ones add their light sample to the path's colour before the bounce is shaded
                    launch_nee(f, sv, nv, d, ln.q[cur], ln.hits, ln.counts, ln.srays, ln.scontrib, c->d_stats + BPT_STAT_RAYS,
                               ln.paths, (unsigned)c->num_sms, ln.st);
                    TraceArgs sa = make_trace_args(c, ln.srays, ln.shits, ln.counts + d, ln.fetch + 2 * (kMaxDepth + 1) + d);
                    sa.count_rays = 0;
                    launch_trace(c, sa, ln.st);
                    launch_nee_resolve(d, ln.counts, ln.shits, ln.scontrib, ln.q[cur].pixel, c->path_color, ln.paths,
                                       (unsigned)c->num_sms, ln.st);
                    c->stats.kernel_launches += 2;
                }
                launch_shade(f, sv, nv, d, ln.q[cur], ln.hits, ln.q[cur ^ 1], ln.counts, ln.fetch, c->path_color, ln.paths,
                             (unsigned)c->num_sms, ln.st);
                c->stats.kernel_launches++;
            }
        for (uint32_t l = 1; l < L; ++l) {
            cudaEventRecord(c->ev_join[l - 1], c->lane_stream[l - 1]);
            cudaStreamWaitEvent(c->stream, c->ev_join[l - 1], 0);
        }
        launch_gather_pass(npix, n, c->path_color, c->frame_sum, c->stream);
        c->stats.kernel_launches++;
    }
    if (!frame_dev) wait_pending_copy(c);  // a pending read-back of the previous frame (inside a capture: done by the caller)
    launch_accumulate(f, frame_dev, c->frame_sum, c->image, c->stream);
    c->stats.kernel_launches++;
}

// Small scenes run the traversal instance that keeps the whole BVH in shared memory (TMA-staged once per CTA).
void plan_staging(bpt_context* c) {
    const uint64_t nodes = (uint64_t)c->blas.num_nodes + (c->two_level ? c->tlas.num_nodes : 0u);
    const uint64_t recs = (uint64_t)c->blas.num_recs + (c->two_level ? c->tlas.num_recs : 0u);
    c->staged = c->built_nodes_ok && nodes <= (uint64_t)c->opt_stage_max_nodes &&
                trace_smem_bytes((uint32_t)recs) <= (size_t)kTraceMaxSmem;
    // the fused path kernel keeps 64 path slots per warp next to the stack: a staged scene must leave room for them
    c->fused_fits = trace_smem_bytes(c->staged ? (uint32_t)recs : 0u, kTraceBlock, true) <= (size_t)kTraceMaxSmem;
}

// inverse of a row-major 3x4 affine transform, in double; false if singular
bool invert3x4(const float* m, float* out) {
    const double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
    const double r = 1.0 / det;
    const double n[9] = {(e * i - f * h) * r, (c * h - b * i) * r, (b * f - c * e) * r,
                         (f * g - d * i) * r, (a * i - c * g) * r, (c * d - a * f) * r,
                         (d * h - e * g) * r, (b * g - a * h) * r, (a * e - b * d) * r};
    const double t[3] = {m[3], m[7], m[11]};
    for (int row = 0; row < 3; ++row) {
        for (int col = 0; col < 3; ++col) out[4 * row + col] = (float)n[3 * row + col];
        out[4 * row + 3] = (float)-(n[3 * row] * t[0] + n[3 * row + 1] * t[1] + n[3 * row + 2] * t[2]);
    }
    for (int k = 0; k < 12; ++k)
        if (!std::isfinite(out[k])) return false;
    return true;
}

}  // namespace

extern "C" {

int bpt_abi_version(void) { return BPT_ABI_VERSION; }

void bpt_params_default(bpt_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->width = 1024; p->height = 1024;          // main.cpp:16-17
    p->spp_per_frame = 32;                      // raygen.rgen:43
    p->max_depth = 8;                           // raygen.rgen:62
    p->frame = 0;
    p->tile_y0 = 0; p->tile_rows = 0;
    p->cam_origin[0] = 0.f; p->cam_origin[1] = -1.f; p->cam_origin[2] = 5.f;   // raygen.rgen:55
    p->cam_target[0] = 0.f; p->cam_target[1] = -1.f; p->cam_target[2] = 2.f;   // raygen.rgen:56
    p->sky[0] = 0.7f; p->sky[1] = 0.6f; p->sky[2] = 0.5f;                      // miss.rmiss:10
    p->tmin = 0.001f; p->tmax = 10000.0f;       // raygen.rgen:71,73
    p->accum_mode = BPT_ACCUM_FLOAT4;
    p->sampler = BPT_SAMPLER_UNIFORM;
}

int bpt_create(int device, void* stream, bpt_context** out) {
    if (!out) return bpt_fail(nullptr, BPT_E_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return bpt_fail(nullptr, BPT_E_CUDA, "no CUDA device available (%s); libbpt has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return bpt_fail(nullptr, BPT_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bpt_fail(nullptr, BPT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bpt_fail(nullptr, BPT_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return bpt_fail(nullptr, BPT_E_CUDA, "device %d is sm_%d%d; libbpt is built for sm_100a only", device, prop.major, prop.minor);
    bpt_context* c = new bpt_context;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if (stream) c->stream = static_cast<cudaStream_t>(stream);
    else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete c;
            return bpt_fail(nullptr, BPT_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    for (int l = 0; l < kMaxLanes - 1 && e == cudaSuccess; ++l) {
        e = cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[l], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_img_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming);
    if (e != cudaSuccess ||
        (e = cudaMalloc(&c->counters, kMaxLanes * kLaneCounters * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&c->d_stats, BPT_STAT_COUNT * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMemsetAsync(c->d_stats, 0, BPT_STAT_COUNT * sizeof(unsigned long long), c->stream)) != cudaSuccess ||
        (e = trace_configure()) != cudaSuccess) {
        int rc = bpt_fail(nullptr, BPT_E_CUDA, "context setup: %s", cudaGetErrorString(e));
        bpt_destroy(c);
        return rc;
    }
    *out = c;
    return BPT_OK;
}

void bpt_destroy(bpt_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int l = 0; l < kMaxLanes - 1; ++l) {
        if (c->lane_stream[l]) { cudaStreamSynchronize(c->lane_stream[l]); cudaStreamDestroy(c->lane_stream[l]); }
        if (c->ev_join[l]) cudaEventDestroy(c->ev_join[l]);
    }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (cudaEvent_t ev : {c->ev_fork, c->ev_img_ready, c->ev_copy_done})
        if (ev) cudaEventDestroy(ev);
    cudaFree(c->d_bgra); cudaFree(c->d_light_prims); cudaFree(c->d_light_cdf);
    if (c->nccl_comm) bpt_nccl_comm_destroy(c->nccl_comm);
    free_scene(c);
    free_paths(c);
    bvh8_free(c->blas); bvh8_free(c->tlas);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    cudaFree(c->d_frame);
    cudaFree(c->d_recs_all); cudaFree(c->d_srec); cudaFree(c->frame_sum);
    cudaFree(c->image); cudaFree(c->image_linear); cudaFree(c->counters); cudaFree(c->d_stats);
    for (auto& p : c->frame_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto& p : c->trace_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* bpt_last_error(const bpt_context* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int bpt_set_option(bpt_context* c, int option, int64_t value) {
    if (!c) return BPT_E_INVALID;
    c->epoch++;  // every option may change what a frame launches
    switch (option) {
        case BPT_OPT_PROFILE: c->profile = value != 0; return BPT_OK;
        case BPT_OPT_COUNT_TRAVERSAL: c->count = value != 0; return BPT_OK;
        case BPT_OPT_SMEM_TOP_NODES:
            if (value < 0) return bpt_fail(c, BPT_E_INVALID, "node count must be >= 0");
            c->opt_stage_max_nodes = value;
            if (c->built) plan_staging(c);
            return BPT_OK;
        case BPT_OPT_TRACE_REFILL_BELOW:
            if (value < 1 || value > 32) return bpt_fail(c, BPT_E_INVALID, "refill threshold must be in [1,32]");
            c->refill_below = (int)value;
            return BPT_OK;
        case BPT_OPT_TRACE_STEPS_PER_REFILL:
            if (value < 1 || value > 64) return bpt_fail(c, BPT_E_INVALID, "steps per refill must be in [1,64]");
            c->steps_per_refill = (int)value;
            return BPT_OK;
        case BPT_OPT_TRACE_STAGED_TRIS_PER_STEP:
            if (value < 1 || value > 16) return bpt_fail(c, BPT_E_INVALID, "triangle tests per step must be in [1,16]");
            c->staged_tris_per_step = (int)value;
            return BPT_OK;
        case BPT_OPT_STREAMS:
            if (value < 1 || value > kMaxLanes) return bpt_fail(c, BPT_E_INVALID, "sample lanes must be in [1,%d]", kMaxLanes);
            c->num_lanes = (int)value;
            return BPT_OK;
        case BPT_OPT_PASS_PATHS:
            if (value < 1) return bpt_fail(c, BPT_E_INVALID, "paths per pass must be >= 1");
            c->pass_paths = value;
            return BPT_OK;
        case BPT_OPT_USE_GRAPH: c->use_graph = value != 0; return BPT_OK;
        case BPT_OPT_FUSED_PATHS: c->fused = value != 0; return BPT_OK;
        case BPT_OPT_BVH_OPTIMAL_COLLAPSE:  // takes effect at the next build of a changed mesh
            c->optimal_collapse = value != 0;
            c->mesh_built = false;
            return BPT_OK;
        case BPT_OPT_BVH_SAH_SUBTREE:  // takes effect at the next build of a changed mesh
            if (value != 0 && (value < 3 || value > 32)) return bpt_fail(c, BPT_E_INVALID, "SAH subtree size must be 0 or in [3,32]");
            c->sah_max_leaves = (uint32_t)value;
            c->mesh_built = false;
            return BPT_OK;
        default: return bpt_fail(c, BPT_E_INVALID, "unknown option %d", option);
    }
}

static int adopt_mesh(bpt_context* c, const void* verts, uint32_t nverts, const void* indices, uint32_t nindices,
                      const void* faces, uint32_t nfaces, cudaMemcpyKind kind) {
    if (!c) return BPT_E_INVALID;
    if (!verts || !indices || !faces) return bpt_fail(c, BPT_E_INVALID, "NULL mesh array");
    if (nindices == 0 || nindices % 3 != 0) return bpt_fail(c, BPT_E_INVALID, "index count %u is not a positive multiple of 3", nindices);
    if (nfaces != nindices / 3) return bpt_fail(c, BPT_E_INVALID, "face count %u != triangle count %u", nfaces, nindices / 3);
    if (nverts == 0) return bpt_fail(c, BPT_E_INVALID, "no vertices");
    cudaSetDevice(c->device);
    if (c->d_verts && c->nverts == nverts && c->nidx == nindices && c->nfaces == nfaces) {
        // same sizes (an animated mesh): keep the allocations, only the contents and the derived structures change
        cudaFree(c->d_xforms); cudaFree(c->d_xforms_inv);
        c->d_xforms = nullptr; c->d_xforms_inv = nullptr;
        c->ninst = 1;
        c->built = false; c->mesh_built = false; c->lights_built = false;
    } else {
        free_scene(c);
        BPT_CUDA_TRY(c, cudaMalloc(&c->d_verts, (size_t)nverts * 12));
        BPT_CUDA_TRY(c, cudaMalloc(&c->d_idx, (size_t)nindices * 4));
        BPT_CUDA_TRY(c, cudaMalloc(&c->d_faces, (size_t)nfaces * 24));
    }
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_verts, verts, (size_t)nverts * 12, kind, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_idx, indices, (size_t)nindices * 4, kind, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_faces, faces, (size_t)nfaces * 24, kind, c->stream));
    // validate the indices on the device copy (one pass at HBM speed instead of a host loop over 3 indices per
    // triangle), then wait: the caller's arrays may be freed after return, like the memcpy at main.cpp:321-325
    uint32_t bad[2] = {0u, 0u};
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->counters, 0, 8, c->stream));
    launch_check_indices(c->d_idx, nindices, nverts, c->counters, c->stream);
    BPT_CUDA_TRY(c, cudaMemcpyAsync(bad, c->counters, 8, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (bad[0]) {
        free_scene(c);
        return bpt_fail(c, BPT_E_INVALID, "%u indices out of range (%u vertices), the first at position %u", bad[0], nverts, bad[1]);
    }
    c->nverts = nverts; c->nidx = nindices; c->nfaces = nfaces; c->ntris = nindices / 3;
    return BPT_OK;
}

int bpt_upload_mesh(bpt_context* c, const float* verts, uint32_t nverts, const uint32_t* indices, uint32_t nindices,
                    const float* faces, uint32_t nfaces) {
    return adopt_mesh(c, verts, nverts, indices, nindices, faces, nfaces, cudaMemcpyHostToDevice);
}
int bpt_upload_mesh_device(bpt_context* c, const void* verts, uint32_t nverts, const void* indices, uint32_t nindices,
                           const void* faces, uint32_t nfaces) {
    return adopt_mesh(c, verts, nverts, indices, nindices, faces, nfaces, cudaMemcpyDeviceToDevice);
}

int bpt_upload_soup(bpt_context* c, uint32_t ntris, uint32_t seed) {
    if (!c) return BPT_E_INVALID;
    if (ntris == 0 || ntris > 0x10000000u) return bpt_fail(c, BPT_E_INVALID, "soup size %u out of range", ntris);
    cudaSetDevice(c->device);
    free_scene(c);
    BPT_CUDA_TRY(c, cudaMalloc(&c->d_verts, (size_t)ntris * 36));
    BPT_CUDA_TRY(c, cudaMalloc(&c->d_idx, (size_t)ntris * 12));
    BPT_CUDA_TRY(c, cudaMalloc(&c->d_faces, (size_t)ntris * 24));
    const float scale = (float)std::pow((double)ntris, -1.0 / 3.0);
    launch_soup(ntris, seed, scale, c->d_verts, c->d_idx, c->d_faces, c->stream);
    BPT_CUDA_TRY(c, cudaGetLastError());
    c->nverts = 3 * ntris; c->nidx = 3 * ntris; c->nfaces = ntris; c->ntris = ntris;
    return BPT_OK;
}

int bpt_upload_obj_arrays(bpt_context* c, const float* positions, uint32_t npositions, const int32_t* corner_vertex, uint32_t ncorners,
                          const int32_t* face_material, const float* materials_kd_ke, uint32_t nmaterials) {
    if (!c) return BPT_E_INVALID;
    if (!positions || !corner_vertex || !face_material || !materials_kd_ke) return bpt_fail(c, BPT_E_INVALID, "NULL scene array");
    if (ncorners == 0 || ncorners % 3 != 0) return bpt_fail(c, BPT_E_INVALID, "corner count %u is not a positive multiple of 3", ncorners);
    if (npositions == 0 || nmaterials == 0) return bpt_fail(c, BPT_E_INVALID, "no positions or no materials");
    cudaSetDevice(c->device);
    free_scene(c);
    const uint32_t nfaces = ncorners / 3;
    float *d_pos = nullptr, *d_mat = nullptr;
    int32_t *d_cv = nullptr, *d_fm = nullptr;
    uint32_t* d_bad = nullptr;
    uint32_t bad = 0;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&d_pos, (size_t)npositions * 12)); step(cudaMalloc(&d_mat, (size_t)nmaterials * 24));
    step(cudaMalloc(&d_cv, (size_t)ncorners * 4)); step(cudaMalloc(&d_fm, (size_t)nfaces * 4)); step(cudaMalloc(&d_bad, 4));
    step(cudaMalloc(&c->d_verts, (size_t)ncorners * 12)); step(cudaMalloc(&c->d_idx, (size_t)ncorners * 4));
    step(cudaMalloc(&c->d_faces, (size_t)nfaces * 24));
    if (e == cudaSuccess) {
        step(cudaMemcpyAsync(d_pos, positions, (size_t)npositions * 12, cudaMemcpyHostToDevice, c->stream));
        step(cudaMemcpyAsync(d_mat, materials_kd_ke, (size_t)nmaterials * 24, cudaMemcpyHostToDevice, c->stream));
        step(cudaMemcpyAsync(d_cv, corner_vertex, (size_t)ncorners * 4, cudaMemcpyHostToDevice, c->stream));
        step(cudaMemcpyAsync(d_fm, face_material, (size_t)nfaces * 4, cudaMemcpyHostToDevice, c->stream));
        step(cudaMemsetAsync(d_bad, 0, 4, c->stream));
        launch_obj_arrays(d_pos, npositions, d_cv, ncorners, d_fm, d_mat, nmaterials, c->d_verts, c->d_idx, c->d_faces, d_bad, c->stream);
        step(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, c->stream));
        step(cudaStreamSynchronize(c->stream));  // the caller's arrays may be freed after return
        step(cudaGetLastError());
    }
    cudaFree(d_pos); cudaFree(d_mat); cudaFree(d_cv); cudaFree(d_fm); cudaFree(d_bad);
    if (e != cudaSuccess) { free_scene(c); return bpt_fail_cuda(c, e, "bpt_upload_obj_arrays", __FILE__, __LINE__); }
    if (bad) { free_scene(c); return bpt_fail(c, BPT_E_INVALID, "%u face corners / faces name a position or material that does not exist", bad); }
    c->nverts = ncorners; c->nidx = ncorners; c->nfaces = nfaces; c->ntris = nfaces;
    return BPT_OK;
}

int bpt_set_instances(bpt_context* c, const float* xforms3x4, uint32_t n) {
    if (!c) return BPT_E_INVALID;
    if (n == 0 || !xforms3x4) return bpt_fail(c, BPT_E_INVALID, "need at least one instance transform");
    if (c->ntris == 0) return bpt_fail(c, BPT_E_STATE, "bpt_set_instances before bpt_upload_mesh");
    if ((uint64_t)n * c->ntris > 0xfffffffeull) return bpt_fail(c, BPT_E_INVALID, "%u instances x %u triangles overflow the 32-bit primitive id", n, c->ntris);
    std::vector<float> inv(12 * (size_t)n);
    for (uint32_t i = 0; i < n; ++i)
        if (!invert3x4(xforms3x4 + 12 * (size_t)i, inv.data() + 12 * (size_t)i))
            return bpt_fail(c, BPT_E_INVALID, "instance %u has a singular or non-finite transform", i);
    cudaSetDevice(c->device);
    cudaFree(c->d_xforms); cudaFree(c->d_xforms_inv);
    c->d_xforms = nullptr; c->d_xforms_inv = nullptr;
    c->built = false;
    c->lights_built = false;
    BPT_CUDA_TRY(c, cudaMalloc(&c->d_xforms, 48 * (size_t)n));
    BPT_CUDA_TRY(c, cudaMalloc(&c->d_xforms_inv, 48 * (size_t)n));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_xforms, xforms3x4, 48 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_xforms_inv, inv.data(), 48 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // inv is a local
    c->ninst = n;
    return BPT_OK;
}

static int build_accel_on_stream(bpt_context* c) {
    // Mesh level (the reference's BLAS, main.cpp:512): only when the mesh changed. Moving the instances
    // (bpt_set_instances + bpt_build_accel) rebuilds just the instance level below, like a TLAS rebuild.
    if (!c->mesh_built) {
        BPT_CUDA_TRY(c, bvh8_alloc(c->blas, c->ntris));
        c->blas.optimal_collapse = c->optimal_collapse;
        c->blas.sah_max_leaves = c->sah_max_leaves;
        bvh8_launch_tri_bounds(c->blas, c->d_verts, c->d_idx, c->stream);
        BPT_CUDA_TRY(c, bvh8_build(c->blas, c->stream));
        if (c->blas.num_leaf_slots != c->ntris)
            return bpt_fail(c, BPT_E_STATE, "BVH8 collapse placed %u of %u triangles", c->blas.num_leaf_slots, c->ntris);
        bvh8_launch_woop(c->blas, c->d_verts, c->d_idx, c->stream);
        if (c->srec_cap != c->ntris) {
            cudaFree(c->d_srec);
            c->d_srec = nullptr; c->srec_cap = 0;
            BPT_CUDA_TRY(c, cudaMalloc(&c->d_srec, (size_t)c->ntris * 64));
            c->srec_cap = c->ntris;
        }
        launch_shade_records(c->d_verts, c->d_idx, c->d_faces, c->ntris, c->d_srec, c->stream);
        c->mesh_built = true;
    }
    c->two_level = c->d_xforms != nullptr;
    cudaFree(c->d_recs_all);
    c->d_recs_all = nullptr;
    uint32_t depth = c->blas.depth;
    if (c->two_level) {
        // K8: the same builder over the instances' world boxes (main.cpp:514-538), then one node array [mesh | instances]
        BPT_CUDA_TRY(c, bvh8_alloc(c->tlas, c->ninst));
        c->tlas.optimal_collapse = c->optimal_collapse;
        c->tlas.sah_max_leaves = c->sah_max_leaves;
        bvh8_launch_instance_bounds(c->tlas, c->d_xforms, c->blas.scene_lo, c->blas.scene_hi, c->stream);
        BPT_CUDA_TRY(c, bvh8_build(c->tlas, c->stream));
        if (c->tlas.num_leaf_slots != c->ninst)
            return bpt_fail(c, BPT_E_STATE, "instance BVH8 collapse placed %u of %u instances", c->tlas.num_leaf_slots, c->ninst);
        bvh8_launch_instance_records(c->tlas, c->d_xforms_inv, c->stream);
        const uint32_t rb = c->blas.num_recs, rt = c->tlas.num_recs;
        BPT_CUDA_TRY(c, cudaMalloc(&c->d_recs_all, (size_t)(rb + rt) * sizeof(Node8)));
        BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_recs_all, c->blas.recs, (size_t)rb * sizeof(Node8), cudaMemcpyDeviceToDevice, c->stream));
        bvh8_launch_append_recs(c->tlas, rb, c->d_recs_all + rb, c->stream);
        depth += c->tlas.depth + 1;  // + the sentinel
    } else {
        bvh8_free(c->tlas);
    }
    if (depth > (uint32_t)kTraceMaxDepth)
        return bpt_fail(c, BPT_E_STATE, "BVH8 depth %u exceeds the traversal stack", depth);
    return BPT_OK;
}

int bpt_build_accel(bpt_context* c) {
    if (!c) return BPT_E_INVALID;
    if (c->ntris == 0) return bpt_fail(c, BPT_E_STATE, "bpt_build_accel before bpt_upload_mesh");
    cudaSetDevice(c->device);
    c->built = false; c->built_nodes_ok = false; c->staged = false;
    c->epoch++;
    cudaEvent_t e0 = get_event(c), e1 = get_event(c);
    cudaEventRecord(e0, c->stream);
    int rc = build_accel_on_stream(c);
    cudaError_t e = cudaEventRecord(e1, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    c->event_pool.push_back(e0); c->event_pool.push_back(e1);  // on every path: the pool owns them
    if (rc != BPT_OK) return rc;
    BPT_CUDA_TRY(c, e);
    c->stats.build_ms = ms;
    c->built_nodes_ok = true;
    plan_staging(c);
    c->built = true;
    return BPT_OK;
}

int bpt_accel_info_get(bpt_context* c, bpt_accel_info* out) {
    if (!c || !out) return BPT_E_INVALID;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "no acceleration structure built");
    memset(out, 0, sizeof(*out));
    out->num_tris = c->ntris;
    out->num_instances = c->ninst;
    out->num_nodes8 = c->blas.num_nodes;
    out->num_binary_nodes = c->ntris - 1;
    out->top_nodes_smem = c->staged ? c->blas.num_recs + (c->two_level ? c->tlas.num_recs : 0u) : 0u;
    out->num_records = c->blas.num_recs;
    out->max_depth8 = c->blas.depth;
    out->num_tlas_nodes8 = c->two_level ? c->tlas.num_nodes : 0;
    out->bytes_nodes = (uint64_t)c->blas.num_nodes * BPT_REC_BYTES;
    out->bytes_tris = (uint64_t)c->ntris * BPT_REC_BYTES;
    return BPT_OK;
}

int bpt_trace(bpt_context* c, const bpt_params* p) {
    if (!c) return BPT_E_INVALID;
    int rc = check_params(c, p);
    if (rc) return rc;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "bpt_trace before bpt_build_accel");
    cudaSetDevice(c->device);
    FrameParams f = to_frame(p);
    const uint32_t npix = tile_local_rows(f) * f.width;
    // samples per pass: enough paths in flight to keep the persistent traversal launches long (their tail and the
    // per-launch fixed costs do not shrink with the tile), whatever the tile size of this GPU
    uint32_t ns = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(f.spp_per_frame, c->pass_paths / npix));
    if ((uint64_t)npix * ns > 0x7fffffffull) ns = 0x7fffffffu / npix;
    if ((rc = ensure_paths(c, (size_t)npix * ns)) != BPT_OK) return rc;
    if ((rc = ensure_frame_sum(c, npix)) != BPT_OK) return rc;
    if ((rc = ensure_image(c, f.width, f.height)) != BPT_OK) return rc;
    if (f.nee) {
        if ((rc = build_light_table(c)) != BPT_OK) return rc;
        if ((rc = ensure_nee(c)) != BPT_OK) return rc;
    }
    if (f.tile_block != c->tile_block || (f.tile_block && (f.tile_nranks != c->tile_nranks || f.tile_rank != c->tile_rank))) {
        // the storage layout of the image buffer changes with the tiling: start from a fresh image
        wait_pending_copy(c);
        BPT_CUDA_TRY(c, cudaMemsetAsync(c->image, 0, (size_t)f.width * f.height * sizeof(float4), c->stream));
        c->tile_block = f.tile_block;
        c->tile_nranks = f.tile_block ? f.tile_nranks : 1;
        c->tile_rank = f.tile_block ? f.tile_rank : 0;
    }
    if (c->frame_events.size() >= kFoldEventsAt) fold_events(c, c->frame_events, c->stats.frame_ms, false);
    if (c->trace_events.size() >= kFoldEventsAt) fold_events(c, c->trace_events, c->stats.trace_kernel_ms, false);
    cudaEvent_t e0 = get_event(c), e1 = get_event(c);
    cudaEventRecord(e0, c->stream);
    // Graph replay (BPT_OPT_USE_GRAPH): frames whose launch list is short enough to be launch-bound (a Cornell box at
    // 256 x 256 is seven launches of a few microseconds each) are captured once and replayed with one graph launch;
    // only the frame index changes between frames, and the kernels read it from d_frame. Per-launch event timing
    // (BPT_OPT_PROFILE) needs individual launches, so it switches the replay off.
    if (c->use_graph && !c->profile) {
        bpt_params key = *p;
        key.frame = 0;
        if (!c->d_frame) BPT_CUDA_TRY(c, cudaMalloc(&c->d_frame, sizeof(int32_t)));
        if (!c->graph_exec || c->graph_epoch != c->epoch || memcmp(&key, &c->graph_key, sizeof(key)) != 0) {
            if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
            const uint64_t k0 = c->stats.kernel_launches, t0 = c->stats.trace_launches;
            cudaGraph_t g = nullptr;
            BPT_CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            enqueue_frame(c, f, npix, ns, c->d_frame);
            BPT_CUDA_TRY(c, cudaStreamEndCapture(c->stream, &g));
            c->graph_kernel_launches = c->stats.kernel_launches - k0;
            c->graph_trace_launches = c->stats.trace_launches - t0;
            c->stats.kernel_launches = k0; c->stats.trace_launches = t0;
            cudaError_t ge = cudaGraphInstantiate(&c->graph_exec, g, 0);
            cudaGraphDestroy(g);
            BPT_CUDA_TRY(c, ge);
            c->graph_key = key;
            c->graph_epoch = c->epoch;
        }
        launch_set_i32(c->d_frame, f.frame, c->stream);
        wait_pending_copy(c);
        BPT_CUDA_TRY(c, cudaGraphLaunch(c->graph_exec, c->stream));
        c->stats.kernel_launches += c->graph_kernel_launches + 1;
        c->stats.trace_launches += c->graph_trace_launches;
    } else {
        enqueue_frame(c, f, npix, ns, nullptr);
    }
    cudaEventRecord(e1, c->stream);
    c->frame_events.emplace_back(e0, e1);
    c->stats.paths += (uint64_t)npix * f.spp_per_frame;
    BPT_CUDA_TRY(c, cudaGetLastError());
    return BPT_OK;
}

int bpt_sync(bpt_context* c) {
    if (!c) return BPT_E_INVALID;
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

int bpt_read_image(bpt_context* c, float* rgba, size_t nfloats) {
    if (!c || !rgba) return BPT_E_INVALID;
    if (!c->image) return bpt_fail(c, BPT_E_STATE, "no image yet");
    size_t need = (size_t)c->img_w * c->img_h * 4;
    if (nfloats < need) return bpt_fail(c, BPT_E_INVALID, "buffer holds %zu floats, image needs %zu", nfloats, need);
    cudaSetDevice(c->device);
    const float4* img = nullptr;
    int rc = row_major_image(c, &img);
    if (rc) return rc;
    BPT_CUDA_TRY(c, cudaMemcpyAsync(rgba, img, need * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

int bpt_read_image_bgra8(bpt_context* c, uint8_t* bgra, size_t nbytes) {
    if (!c || !bgra) return BPT_E_INVALID;
    if (!c->image) return bpt_fail(c, BPT_E_STATE, "no image yet");
    size_t npix = (size_t)c->img_w * c->img_h;
    if (nbytes < npix * 4) return bpt_fail(c, BPT_E_INVALID, "buffer holds %zu bytes, image needs %zu", nbytes, npix * 4);
    cudaSetDevice(c->device);
    const float4* img = nullptr;
    int rc = row_major_image(c, &img);
    if (rc) return rc;
    wait_pending_copy(c);
    launch_image_to_bgra8(img, c->d_bgra, npix, c->stream);
    BPT_CUDA_TRY(c, cudaMemcpyAsync(bgra, c->d_bgra, npix * 4, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

// Asynchronous read-back: the copy runs on the context's copy stream behind an event of the main stream, so the next
// frame's kernels overlap it; whoever writes the image next waits for ev_copy_done on the device (wait_pending_copy).
static int read_async(bpt_context* c, void* host, size_t have, bool bgra8) {
    if (!c || !host) return BPT_E_INVALID;
    if (!c->image) return bpt_fail(c, BPT_E_STATE, "no image yet");
    const size_t npix = (size_t)c->img_w * c->img_h;
    const size_t need = bgra8 ? npix * 4 : npix * 4 * sizeof(float);
    if (have < need) return bpt_fail(c, BPT_E_INVALID, "buffer holds %zu bytes, image needs %zu", have, need);
    cudaSetDevice(c->device);
    const float4* img = nullptr;
    int rc = row_major_image(c, &img);
    if (rc) return rc;
    const void* src = img;
    if (bgra8) {
        wait_pending_copy(c);
        launch_image_to_bgra8(img, c->d_bgra, npix, c->stream);
        src = c->d_bgra;
    }
    BPT_CUDA_TRY(c, cudaEventRecord(c->ev_img_ready, c->stream));
    BPT_CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_img_ready, 0));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(host, src, need, cudaMemcpyDeviceToHost, c->copy_stream));
    BPT_CUDA_TRY(c, cudaEventRecord(c->ev_copy_done, c->copy_stream));
    c->copy_pending = true;
    return BPT_OK;
}
int bpt_read_image_async(bpt_context* c, float* rgba, size_t nfloats) { return read_async(c, rgba, nfloats * sizeof(float), false); }
int bpt_read_image_bgra8_async(bpt_context* c, uint8_t* bgra, size_t nbytes) { return read_async(c, bgra, nbytes, true); }
int bpt_read_wait(bpt_context* c) {
    if (!c) return BPT_E_INVALID;
    if (!c->copy_pending) return BPT_OK;
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaEventSynchronize(c->ev_copy_done));
    c->copy_pending = false;
    return BPT_OK;
}

int bpt_image_device_ptr(bpt_context* c, void** dptr, size_t* nbytes) {
    if (!c || !dptr) return BPT_E_INVALID;
    if (!c->image) return bpt_fail(c, BPT_E_STATE, "no image yet");
    cudaSetDevice(c->device);
    const float4* img = nullptr;
    int rc = row_major_image(c, &img);  // stream-ordered: valid for work enqueued after this call
    if (rc) return rc;
    *dptr = const_cast<float4*>(img);
    if (nbytes) *nbytes = (size_t)c->img_w * c->img_h * sizeof(float4);
    return BPT_OK;
}

int bpt_clear_image(bpt_context* c) {
    if (!c) return BPT_E_INVALID;
    cudaSetDevice(c->device);
    wait_pending_copy(c);
    if (c->image) BPT_CUDA_TRY(c, cudaMemsetAsync(c->image, 0, (size_t)c->img_w * c->img_h * sizeof(float4), c->stream));
    if (c->frame_sum) BPT_CUDA_TRY(c, cudaMemsetAsync(c->frame_sum, 0, c->cap_pixels * sizeof(float4), c->stream));
    return BPT_OK;
}

int bpt_get_stats(bpt_context* c, bpt_stats* out) {
    if (!c || !out) return BPT_E_INVALID;
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned long long h[BPT_STAT_COUNT];
    BPT_CUDA_TRY(c, cudaMemcpy(h, c->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    c->stats.rays_traced = h[BPT_STAT_RAYS];
    c->stats.nodes_visited = h[BPT_STAT_NODES];
    c->stats.tris_tested = h[BPT_STAT_TRIS];
    c->stats.warp_iterations = h[BPT_STAT_WARP_ITERS];
    c->stats.warp_node_steps = h[BPT_STAT_WARP_NODE_STEPS];
    c->stats.warp_tri_steps = h[BPT_STAT_WARP_TRI_STEPS];
    c->stats.lane_iterations = h[BPT_STAT_LANE_ITERS];
    fold_events(c, c->frame_events, c->stats.frame_ms, true);
    fold_events(c, c->trace_events, c->stats.trace_kernel_ms, true);
    *out = c->stats;
    return BPT_OK;
}

int bpt_reset_stats(bpt_context* c) {
    if (!c) return BPT_E_INVALID;
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    bpt_stats tmp;
    bpt_get_stats(c, &tmp);
    double build = c->stats.build_ms;
    c->stats = bpt_stats{};
    c->stats.build_ms = build;
    BPT_CUDA_TRY(c, cudaMemset(c->d_stats, 0, BPT_STAT_COUNT * sizeof(unsigned long long)));
    return BPT_OK;
}

int bpt_trace_rays(bpt_context* c, const float* rays, uint32_t n, void* hits) {
    if (!c || !rays || !hits) return BPT_E_INVALID;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "bpt_trace_rays before bpt_build_accel");
    if (n == 0) return BPT_OK;
    cudaSetDevice(c->device);
    int rc = ensure_paths(c, n);
    if (rc) return rc;
    uint32_t init[2] = {n, 0u};
    uint32_t* counts = c->counters;
    uint32_t* fetch = c->counters + (kMaxDepth + 1);
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->q[0].rays, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(counts, &init[0], 4, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(fetch, &init[1], 4, cudaMemcpyHostToDevice, c->stream));
    launch_trace(c, make_trace_args(c, c->q[0].rays, c->hits, counts, fetch), c->stream);
    launch_refine_hits(SceneView{c->d_srec, c->d_xforms, c->ntris}, c->q[0].rays, c->hits, n, c->stream);
    BPT_CUDA_TRY(c, cudaMemcpyAsync(hits, c->hits, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    BPT_CUDA_TRY(c, cudaGetLastError());
    return BPT_OK;
}

int bpt_generate_rays(bpt_context* c, const bpt_params* p, uint32_t sample_in_frame, float* rays, uint32_t* seeds) {
    if (!c || !rays || !seeds) return BPT_E_INVALID;
    int rc = check_params(c, p);
    if (rc) return rc;
    cudaSetDevice(c->device);
    FrameParams f = to_frame(p);
    const uint32_t npix = tile_local_rows(f) * f.width;
    if ((rc = ensure_paths(c, npix)) != BPT_OK) return rc;
    uint32_t* counts = c->counters;
    uint32_t* fetch = c->counters + (kMaxDepth + 1);
    launch_generate(f, nullptr, sample_in_frame, 1, 0u, c->q[0], counts, fetch, f.max_depth + 1, c->stream);
    std::vector<float4> st(npix);
    BPT_CUDA_TRY(c, cudaMemcpyAsync(rays, c->q[0].rays, (size_t)npix * 32, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(st.data(), c->q[0].state, (size_t)npix * 16, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < npix; ++i) memcpy(&seeds[i], &st[i].w, 4);
    return BPT_OK;
}

int bpt_shade_step(bpt_context* c, const bpt_params* p, uint32_t n, const float* rays, const void* hits, const float* weight,
                   const uint32_t* seed, float* contrib, float* new_rays, float* new_weight, uint32_t* new_seed, uint8_t* alive) {
    if (!c || !rays || !hits || !weight || !seed || !contrib || !new_rays || !new_weight || !new_seed || !alive) return BPT_E_INVALID;
    int rc = check_params(c, p);
    if (rc) return rc;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "bpt_shade_step before bpt_build_accel");
    if (n == 0) return BPT_OK;
    cudaSetDevice(c->device);
    if ((rc = ensure_paths(c, n)) != BPT_OK) return rc;
    FrameParams f = to_frame(p);
    f.max_depth = 2;  // shade bounce 0 of 2: the next segment is sampled, as raygen.rgen:78-80 always does
    f.nee = 0;        // a single step has no previous vertex: camera-ray semantics (emission is added in full)
    std::vector<float4> st(n);
    std::vector<uint32_t> pix(n);
    for (uint32_t i = 0; i < n; ++i) {
        st[i] = make_float4(weight[3 * (size_t)i], weight[3 * (size_t)i + 1], weight[3 * (size_t)i + 2], 0.f);
        memcpy(&st[i].w, &seed[i], 4);
        pix[i] = i;
    }
    uint32_t* counts = c->counters;
    uint32_t* fetch = c->counters + (kMaxDepth + 1);
    const uint32_t init[2] = {n, 0u};
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->q[0].rays, rays, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->q[0].state, st.data(), (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->q[0].pixel, pix.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->hits, hits, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(counts, init, 8, cudaMemcpyHostToDevice, c->stream));          // counts[0] = n, counts[1] = 0
    BPT_CUDA_TRY(c, cudaMemcpyAsync(fetch + kCounterStride, init + 1, 4, cudaMemcpyHostToDevice, c->stream));  // tile counter of bounce 0
    launch_shade(f, SceneView{c->d_srec, c->d_xforms, c->ntris}, NeeView{nullptr, nullptr, 0u, 0.f, nullptr}, 0u, c->q[0], c->hits,
                 c->q[1], counts, fetch, c->path_color, n, (unsigned)c->num_sms, c->stream);
    uint32_t m = 0;
    std::vector<float4> col(n), orays(2 * (size_t)n), ost(n);
    BPT_CUDA_TRY(c, cudaMemcpyAsync(&m, counts + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaMemcpyAsync(col.data(), c->path_color, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->path_color, 0, (size_t)n * 16, c->stream));  // the per-path colours are zero between passes
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    BPT_CUDA_TRY(c, cudaGetLastError());
    if (m > n) return bpt_fail(c, BPT_E_STATE, "shade step produced %u survivors of %u paths", m, n);
    if (m) {
        BPT_CUDA_TRY(c, cudaMemcpy(orays.data(), c->q[1].rays, (size_t)m * 32, cudaMemcpyDeviceToHost));
        BPT_CUDA_TRY(c, cudaMemcpy(ost.data(), c->q[1].state, (size_t)m * 16, cudaMemcpyDeviceToHost));
        BPT_CUDA_TRY(c, cudaMemcpy(pix.data(), c->q[1].pixel, (size_t)m * 4, cudaMemcpyDeviceToHost));
    }
    memset(new_rays, 0, (size_t)n * 32); memset(new_weight, 0, (size_t)n * 12); memset(new_seed, 0, (size_t)n * 4); memset(alive, 0, n);
    for (uint32_t i = 0; i < n; ++i) { contrib[3 * (size_t)i] = col[i].x; contrib[3 * (size_t)i + 1] = col[i].y; contrib[3 * (size_t)i + 2] = col[i].z; }
    for (uint32_t j = 0; j < m; ++j) {  // the survivors were compacted: scatter them back to their path
        const uint32_t i = pix[j];
        if (i >= n) return bpt_fail(c, BPT_E_STATE, "shade step returned path id %u of %u", i, n);
        memcpy(new_rays + 8 * (size_t)i, &orays[2 * (size_t)j], 32);
        memcpy(new_weight + 3 * (size_t)i, &ost[j], 12);
        memcpy(new_seed + i, &ost[j].w, 4);
        alive[i] = 1;
    }
    return BPT_OK;
}

int bpt_download_accel(bpt_context* c, void* records, uint32_t* rec_prim, float* grid) {
    if (!c) return BPT_E_INVALID;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "no acceleration structure built");
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (records) BPT_CUDA_TRY(c, cudaMemcpy(records, c->blas.recs, (size_t)c->blas.num_recs * BPT_REC_BYTES, cudaMemcpyDeviceToHost));
    if (rec_prim) BPT_CUDA_TRY(c, cudaMemcpy(rec_prim, c->blas.rec_prim, (size_t)c->blas.num_recs * 4, cudaMemcpyDeviceToHost));
    if (grid)
        for (int k = 0; k < 3; ++k) { grid[k] = c->blas.grid_bias[k]; grid[3 + k] = c->blas.grid_step[k]; }
    return BPT_OK;
}

int bpt_download_mesh(bpt_context* c, float* verts, uint32_t* indices, float* faces) {
    if (!c) return BPT_E_INVALID;
    if (c->ntris == 0) return bpt_fail(c, BPT_E_STATE, "no mesh uploaded");
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (verts) BPT_CUDA_TRY(c, cudaMemcpy(verts, c->d_verts, (size_t)c->nverts * 12, cudaMemcpyDeviceToHost));
    if (indices) BPT_CUDA_TRY(c, cudaMemcpy(indices, c->d_idx, (size_t)c->nidx * 4, cudaMemcpyDeviceToHost));
    if (faces) BPT_CUDA_TRY(c, cudaMemcpy(faces, c->d_faces, (size_t)c->nfaces * 24, cudaMemcpyDeviceToHost));
    return BPT_OK;
}

int bpt_download_morton(bpt_context* c, uint64_t* keys, uint32_t n) {
    if (!c || !keys) return BPT_E_INVALID;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "no acceleration structure built");
    if (n != c->ntris) return bpt_fail(c, BPT_E_INVALID, "n must equal the triangle count %u", c->ntris);
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    BPT_CUDA_TRY(c, cudaMemcpy(keys, c->blas.keys, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return BPT_OK;
}

int bpt_download_leaf_order(bpt_context* c, uint32_t* prims, uint32_t n) {
    if (!c || !prims) return BPT_E_INVALID;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "no acceleration structure built");
    if (n != c->ntris) return bpt_fail(c, BPT_E_INVALID, "n must equal the triangle count %u", c->ntris);
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    BPT_CUDA_TRY(c, cudaMemcpy(prims, c->blas.leaf_prim, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return BPT_OK;
}

int bpt_download_lbvh(bpt_context* c, uint32_t* left, uint32_t* right, float* aabbs) {
    if (!c) return BPT_E_INVALID;
    if (!c->built) return bpt_fail(c, BPT_E_STATE, "no acceleration structure built");
    cudaSetDevice(c->device);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    const uint32_t n = c->ntris;
    if (left && n > 1) BPT_CUDA_TRY(c, cudaMemcpy(left, c->blas.left, (size_t)(n - 1) * 4, cudaMemcpyDeviceToHost));
    if (right && n > 1) BPT_CUDA_TRY(c, cudaMemcpy(right, c->blas.right, (size_t)(n - 1) * 4, cudaMemcpyDeviceToHost));
    if (aabbs) {
        std::vector<float4> lo(2 * (size_t)n - 1), hi(2 * (size_t)n - 1);
        BPT_CUDA_TRY(c, cudaMemcpy(lo.data(), c->blas.nlo, lo.size() * 16, cudaMemcpyDeviceToHost));
        BPT_CUDA_TRY(c, cudaMemcpy(hi.data(), c->blas.nhi, hi.size() * 16, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < lo.size(); ++i) {
            aabbs[6 * i + 0] = lo[i].x; aabbs[6 * i + 1] = lo[i].y; aabbs[6 * i + 2] = lo[i].z;
            aabbs[6 * i + 3] = hi[i].x; aabbs[6 * i + 4] = hi[i].y; aabbs[6 * i + 5] = hi[i].z;
        }
    }
    return BPT_OK;
}

// ---------------------------------------------------------------- multi-GPU
void bpt_tile_rows(uint32_t height, int rank, int nranks, uint32_t* y0, uint32_t* rows) {
    if (nranks < 1) nranks = 1;
    uint32_t per = height / (uint32_t)nranks;
    if (y0) *y0 = per * (uint32_t)rank;
    if (rows) *rows = rank == nranks - 1 ? height - per * (uint32_t)rank : per;
}

int bpt_nccl_unique_id(uint8_t id[BPT_NCCL_UNIQUE_ID_BYTES]) {
    std::string err;
    if (bpt_nccl_get_unique_id(id, &err) != 0) return bpt_fail(nullptr, BPT_E_NCCL, "%s", err.c_str());
    return BPT_OK;
}

int bpt_nccl_init(bpt_context* c, const uint8_t id[BPT_NCCL_UNIQUE_ID_BYTES], int rank, int nranks) {
    if (!c || !id) return BPT_E_INVALID;
    if (nranks < 1 || rank < 0 || rank >= nranks) return bpt_fail(c, BPT_E_INVALID, "bad rank %d of %d", rank, nranks);
    cudaSetDevice(c->device);
    if (c->nccl_comm) { bpt_nccl_comm_destroy(c->nccl_comm); c->nccl_comm = nullptr; }
    std::string err;
    if (bpt_nccl_comm_init(&c->nccl_comm, id, rank, nranks, &err) != 0) return bpt_fail(c, BPT_E_NCCL, "%s", err.c_str());
    c->rank = rank;
    c->nranks = nranks;
    return BPT_OK;
}

int bpt_allgather_image(bpt_context* c, uint32_t width, uint32_t height) {
    if (!c) return BPT_E_INVALID;
    if (!c->nccl_comm) return bpt_fail(c, BPT_E_STATE, "bpt_allgather_image before bpt_nccl_init");
    if (!c->image || c->img_w != width || c->img_h != height) return bpt_fail(c, BPT_E_STATE, "image is not %ux%u", width, height);
    if (height % (uint32_t)c->nranks) return bpt_fail(c, BPT_E_INVALID, "height %u not divisible by %d ranks", height, c->nranks);
    if (c->tile_block && ((int)c->tile_nranks != c->nranks || (int)c->tile_rank != c->rank))
        return bpt_fail(c, BPT_E_STATE, "interleaved tile %u/%u does not match NCCL rank %d/%d", c->tile_rank, c->tile_nranks, c->rank, c->nranks);
    cudaSetDevice(c->device);
    // contiguous and interleaved tilings both keep rank r's rows at row r*height/nranks of the buffer
    uint32_t y0, rows;
    bpt_tile_rows(height, c->rank, c->nranks, &y0, &rows);
    wait_pending_copy(c);  // the gather overwrites the other ranks' rows a pending read-back may still be copying
    size_t count = (size_t)rows * width * 4;  // floats per rank
    std::string err;
    // in place: this rank's tile already sits at its slot of the buffer (K12 wrote it there)
    if (bpt_nccl_allgather_f32(c->nccl_comm, c->image + (size_t)y0 * width, c->image, count, c->stream, &err) != 0)
        return bpt_fail(c, BPT_E_NCCL, "%s", err.c_str());
    return BPT_OK;
}

}  // extern "C"
