// shade.cu — K9 generate, K11 shade, K12 accumulate (+ the synthetic soup generator).
//
// This translation unit is compiled with --fmad=false and without fast-math so that every
// float operation is the single IEEE operation the shader text names, in the same order as the
// CPU checker (tests) restates it; the only library calls whose last bit may differ from a host libm are
// sinf/cosf. Reference lines (paths relative to the reference checkout):
//   K9  shaders/raygen.rgen:47-60   seed, jitter, camera ray, weight = 1
//   K11 shaders/closesthit.rchit:50-64, shaders/miss.rmiss:8-12, shaders/raygen.rgen:14-39,76-83
//   K12 shaders/raygen.rgen:86-90   (rgba8 emulation: main.cpp:481-484)
#include <algorithm>

#include "shade_one.cuh"

namespace {

constexpr int kBlock = 256;
constexpr unsigned FULL = 0xffffffffu;

using namespace bpt_shade;


// ---------------------------------------------------------------- K9
// frame_dev: when non-null, the frame index is read from device memory instead of p.frame, so that a captured CUDA
// graph of a frame's launches can be replayed for every frame (api.cu, BPT_OPT_USE_GRAPH).
__global__ void k_generate(FrameParams p, const int32_t* __restrict__ frame_dev, uint32_t s0, uint32_t ns, uint32_t path_base,
                           PathQueue q, uint32_t* counts, uint32_t* fetch, uint32_t ncounters) {
    if (frame_dev) p.frame = *frame_dev;
    const uint32_t npix = tile_local_rows(p) * p.width;
    const uint32_t npaths = npix * ns;  // one pass carries samples s0 .. s0+ns-1 of every tile pixel
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncounters) {  // queue lengths and fetch counters of this sample pass
        counts[i] = i == 0 ? npaths : 0u;
        fetch[i] = 0u;
        fetch[kCounterStride + i] = 0u;      // tile counters of k_shade
        fetch[2 * kCounterStride + i] = 0u;  // fetch counters of the shadow-ray launches
    }
    if (i >= npaths) return;
    float4 ro, rd;
    uint32_t seed;
    gen_primary(p, s0, npix, i, ro, rd, seed);
    q.rays[2 * (size_t)i] = ro;
    q.rays[2 * (size_t)i + 1] = rd;
    q.state[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed));
    q.pixel[i] = path_base + i;  // path id of the pass: (sample slot of the pass) * npix + tile-local pixel
}

// ---------------------------------------------------------------- K11
// Shading record of a primitive (64 bytes, 32-byte aligned, built once by k_shade_records): the three vertices the
// index buffer names (closesthit.rchit:52-54) and its Face {Kd, Ke} (:60-62), so that shading gathers two sectors
// instead of an index triple, three vertices and a face record.
struct ShadeRec { V3 v0, v1, v2, kd, ke; };
__device__ __forceinline__ ShadeRec load_rec(const SceneView& s, uint32_t prim, const float* m) {
    const float4* r = s.srec + 4 * (size_t)prim;
    const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2), d = __ldg(r + 3);
    return {xform(m, V3{a.x, a.y, a.z}), xform(m, V3{a.w, b.x, b.y}), xform(m, V3{b.z, b.w, c.x}), V3{c.y, c.z, c.w},
            V3{d.x, d.y, d.z}};
}
__global__ void k_shade_records(const float* __restrict__ verts, const uint32_t* __restrict__ idx,
                                const float* __restrict__ faces, uint32_t n, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float f[16];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* v = verts + 3 * (size_t)idx[3 * (size_t)i + c];
        f[3 * c] = v[0]; f[3 * c + 1] = v[1]; f[3 * c + 2] = v[2];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) f[9 + k] = faces[6 * (size_t)i + k];
    f[15] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) out[4 * (size_t)i + q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
}


// bpt_trace_rays: the same refinement applied to a hit buffer, so stage-level callers see what shading sees.
__global__ void k_refine_hits(SceneView s, const float4* __restrict__ rays, uint4* __restrict__ hits, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 h = hits[i];
    if (h.w == BPT_MISS) return;
    const uint32_t inst = s.xforms ? h.w / s.ntris : 0u;
    const uint32_t prim = s.xforms ? h.w - inst * s.ntris : h.w;
    const float* m = s.xforms ? s.xforms + 12 * (size_t)inst : nullptr;
    const ShadeRec sr = load_rec(s, prim, m);
    const V3 v0 = sr.v0, v1 = sr.v1, v2 = sr.v2;
    const float4 ro = rays[2 * (size_t)i], rd = rays[2 * (size_t)i + 1];
    float u, v, t = __uint_as_float(h.x);
    barycentrics(V3{ro.x, ro.y, ro.z}, V3{rd.x, rd.y, rd.z}, v0, v1, v2, u, v, t);
    h.x = __float_as_uint(t); h.y = __float_as_uint(u); h.z = __float_as_uint(v);
    hits[i] = h;
}



// warp-ballot compaction of the surviving paths into the next queue: one atomic per warp
__device__ __forceinline__ void compact_out(bool alive, const ShadeOut& o, uint32_t pix, PathQueue out, uint32_t* count_next,
                                            unsigned lane) {
    const unsigned live = __ballot_sync(FULL, alive);
    if (!live) return;
    uint32_t base = 0;
    if (lane == (unsigned)(__ffs(live) - 1)) base = atomicAdd(count_next, (uint32_t)__popc(live));
    base = __shfl_sync(FULL, base, __ffs(live) - 1);
    if (alive) {
        const uint32_t j = base + __popc(live & ((1u << lane) - 1u));
        out.rays[2 * (size_t)j] = o.ro;
        out.rays[2 * (size_t)j + 1] = o.rd;
        out.state[j] = o.st;
        out.pixel[j] = pix;
    }
}

// One 256-path tile at a time per block, everything loaded where it is used: a dependent chain tile counter -> hit ->
// shading record -> stores, hidden by 32 resident warps per SM. ncu (profiles/r2b_k_shade_plain_ncu_soup10m.txt):
// 14.4 GB of DRAM traffic in 3.26 ms = 4.4 TB/s = 67 % of the measured copy bandwidth (54 % of ncu's theoretical peak),
// issue slots 44 % busy (IEEE division / square root sequences). A cp.async ring instance (three stages per warp, every
// load of a path in flight while others are shaded, no barriers) was built and measured in round 2: with the 50 KB of
// shared memory per 128-thread block it runs 16 warps per SM instead of 32, which the arithmetic of a shade step
// needs more than the loads need the ring: 3.56 ms per launch on the soup, and 4 x slower on the Cornell box
// (profiles/r2d_*); removed again.
template <bool EXTRA>
__global__ void __launch_bounds__(kBlock, 4) k_shade(FrameParams p, SceneView s, uint32_t depth, PathQueue in, const uint4* __restrict__ hits,
                        PathQueue out, uint32_t* counts, uint32_t* tile_ctr, float4* path_color, float* pdf_prev,
                        float light_area) {
    // The host does not know how many paths are still alive, so a grid of a few waves per SM pulls 256-path tiles
    // from a counter, in order (one block per 256 slots of the FULL queue would launch and retire half a million
    // mostly empty blocks per bounce; a strided or chunked loop would interleave distant paths in the compacted
    // output and cost the traversal kernel its ray coherence — measured: -4 %).
    const uint32_t n = counts[depth];
    const unsigned lane = threadIdx.x & 31u;
    __shared__ uint32_t s_tile;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_ctr, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        __syncthreads();
        if ((uint64_t)tile * blockDim.x >= n) break;
        const uint32_t i = tile * blockDim.x + threadIdx.x;
        bool alive = false;
        ShadeOut o;
        uint32_t pix = 0;
        if (i < n) {
            const uint4 h = hits[i];
            const float4 st = in.state[i];
            pix = in.pixel[i];
            float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = ro, ra = ro, rb = ro, rc = ro, rdd = ro;
            if (h.w != BPT_MISS) {
                const float4* r = s.srec + 4 * (size_t)(s.xforms ? h.w % s.ntris : h.w);
                ra = __ldg(r); rb = __ldg(r + 1); rc = __ldg(r + 2); rdd = __ldg(r + 3);
                ro = in.rays[2 * (size_t)i]; rd = in.rays[2 * (size_t)i + 1];
            }
            alive = shade_one<EXTRA>(p, s, depth, h, st, pix, ro, rd, ra, rb, rc, rdd, path_color, pdf_prev, light_area, o);
        }
        compact_out(alive, o, pix, out, &counts[depth + 1], lane);
    }
}

// ---------------------------------------------------------------- next-event estimation (bpt.h; not the reference)
// One thread per path of bounce `depth`'s queue; the formulas are the ones the CPU checker states for the same switch,
// in the same order, so that both agree with the same seeds.
__global__ void __launch_bounds__(kBlock) k_nee(FrameParams p, SceneView s, NeeView nv, uint32_t depth, PathQueue in,
                                                const uint4* __restrict__ hits, const uint32_t* __restrict__ counts,
                                                float4* __restrict__ shadow_rays, float4* __restrict__ shadow_contrib,
                                                unsigned long long* ray_stat) {
    const uint32_t n = counts[depth];
    uint32_t cast = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 so = make_float4(0.f, 0.f, 0.f, p.tmin), sd = make_float4(0.f, 0.f, 1.f, -1.0f);  // tmax < tmin: no connection
        float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint4 h = hits[i];
        if (h.w != BPT_MISS && depth + 1u < p.max_depth) {
            const float4 st = in.state[i];
            uint32_t seed = __float_as_uint(st.w);
            const V3 w{st.x, st.y, st.z};
            // instanced scenes: world primitive id = instance * ntris + triangle, vertices through the instance matrix
            const ShadeRec sr = load_rec(s, s.xforms ? h.w % s.ntris : h.w, s.xforms ? s.xforms + 12 * (size_t)(h.w / s.ntris) : nullptr);
            const float4 ro = in.rays[2 * (size_t)i], rd = in.rays[2 * (size_t)i + 1];
            float u, v, t = __uint_as_float(h.x);
            barycentrics(V3{ro.x, ro.y, ro.z}, V3{rd.x, rd.y, rd.z}, sr.v0, sr.v1, sr.v2, u, v, t);
            const float b0 = 1.0f - u - v;
            const V3 pos = sr.v0 * b0 + sr.v1 * u + sr.v2 * v;
            const V3 nrm = -normalize(cross(sr.v1 - sr.v0, sr.v2 - sr.v0));
            const V3 brdf = sr.kd / kPi;
            const float rs = bpt_rand(seed), ra = bpt_rand(seed), rb = bpt_rand(seed);
            in.state[i].w = __uint_as_float(seed);
            uint32_t lo = 0, hi = nv.nlights - 1u;  // first entry with cdf > rs (the last one if rs == 1)
            while (lo < hi) {
                const uint32_t mid = (lo + hi) / 2u;
                if (__ldg(nv.light_cdf + mid) > rs) hi = mid; else lo = mid + 1u;
            }
            const uint32_t lp = __ldg(nv.light_prims + lo);
            const ShadeRec lr = load_rec(s, s.xforms ? lp % s.ntris : lp, s.xforms ? s.xforms + 12 * (size_t)(lp / s.ntris) : nullptr);
            const float su = sqrtf(ra);
            const float bu = su * (1.0f - rb), bv = su * rb, lb0 = 1.0f - bu - bv;
            const V3 y = lr.v0 * lb0 + lr.v1 * bu + lr.v2 * bv;
            const V3 l = y - pos;
            const float r2l = dot(l, l);
            if (lp != h.w && r2l > 0.0f) {
                const float rl = sqrtf(r2l);
                const V3 wd = l / rl;
                const V3 ny = -normalize(cross(lr.v1 - lr.v0, lr.v2 - lr.v0));
                const float cx = dot(wd, nrm), cy = fabsf(dot(wd, ny));
                if (cx > 0.0f && cy > 0.0f) {
                    const float pl = r2l / (cy * nv.light_area);
                    const float pb = p.sampler == BPT_SAMPLER_COSINE ? cx / kPi : kPdf;
                    const V3 c = w * brdf * lr.ke * (cx / (pl + pb));
                    so = make_float4(pos.x, pos.y, pos.z, p.tmin);
                    sd = make_float4(wd.x, wd.y, wd.z, rl * 0.999f);
                    sc = make_float4(c.x, c.y, c.z, 1.0f);
                    ++cast;
                }
            }
        }
        shadow_rays[2 * (size_t)i] = so;
        shadow_rays[2 * (size_t)i + 1] = sd;
        shadow_contrib[i] = sc;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cast += __shfl_xor_sync(FULL, cast, o);
    if ((threadIdx.x & 31u) == 0u && cast) atomicAdd(ray_stat, (unsigned long long)cast);
}

__global__ void k_nee_resolve(uint32_t depth, const uint32_t* __restrict__ counts, const uint4* __restrict__ shadow_hits,
                              const float4* __restrict__ shadow_contrib, const uint32_t* __restrict__ pixel,
                              float4* __restrict__ path_color) {
    const uint32_t n = counts[depth];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 c = shadow_contrib[i];
        if (c.w == 0.0f || shadow_hits[i].w != BPT_MISS) continue;  // no connection, or something is in the way
        const uint32_t pix = pixel[i];
        float4 acc = path_color[pix];
        acc.x += c.x; acc.y += c.y; acc.z += c.z;
        path_color[pix] = acc;
    }
}

// frame_sum[pixel] += color of samples s0..s0+ns-1, in sample order (raygen.rgen:76 adds into one `color` per pixel)
__global__ void k_gather_pass(uint32_t npix, uint32_t ns, float4* __restrict__ path_color, float4* __restrict__ frame_sum) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 acc = frame_sum[i];
    for (uint32_t s0 = 0; s0 < ns; s0 += 8) {  // 8 independent loads in flight, then the adds in sample order
        float4 c[8];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k)
            c[k] = s0 + k < ns ? path_color[(size_t)(s0 + k) * npix + i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            if (s0 + k < ns) {
                path_color[(size_t)(s0 + k) * npix + i] = make_float4(0.f, 0.f, 0.f, 0.f);
                acc.x += c[k].x; acc.y += c[k].y; acc.z += c[k].z;
            }
        }
    }
    frame_sum[i] = acc;
}

// ---------------------------------------------------------------- K12
__device__ __forceinline__ float unorm8_roundtrip(float x) {
    float c = fminf(fmaxf(x, 0.0f), 1.0f);  // NaN -> 0 (fmaxf drops the NaN)
    return rintf(c * 255.0f) / 255.0f;
}
__global__ void k_accumulate(FrameParams p, const int32_t* __restrict__ frame_dev, float4* __restrict__ frame_sum,
                             float4* __restrict__ image) {
    if (frame_dev) p.frame = *frame_dev;
    const uint32_t npix = tile_local_rows(p) * p.width;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 c = frame_sum[i];
    frame_sum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float spp = (float)p.spp_per_frame;
    c.x = c.x / spp; c.y = c.y / spp; c.z = c.z / spp;               // raygen.rgen:86
    float4* dst = image + (size_t)tile_store_row(p, 0) * p.width + i;  // the tile's rows are contiguous in the buffer
    const float4 old = *dst;                                         // :88
    const float fr = (float)p.frame, fr1 = (float)(p.frame + 1);
    float4 nw;                                                       // :89
    nw.x = (c.x + old.x * fr) / fr1;
    nw.y = (c.y + old.y * fr) / fr1;
    nw.z = (c.z + old.z * fr) / fr1;
    nw.w = (1.0f + old.w * fr) / fr1;
    if (p.accum_mode == BPT_ACCUM_RGBA8) {
        nw.x = unorm8_roundtrip(nw.x); nw.y = unorm8_roundtrip(nw.y);
        nw.z = unorm8_roundtrip(nw.z); nw.w = unorm8_roundtrip(nw.w);
    }
    *dst = nw;                                                       // :90
}

__global__ void k_image_to_bgra8(const float4* __restrict__ image, uint8_t* __restrict__ bgra, size_t npix) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 c = image[i];
    auto q = [](float x) { return (uint8_t)rintf(fminf(fmaxf(x, 0.0f), 1.0f) * 255.0f); };
    reinterpret_cast<uchar4*>(bgra)[i] = make_uchar4(q(c.z), q(c.y), q(c.x), q(c.w));  // B8G8R8A8 (main.cpp:483)
}

__global__ void k_deinterleave(const float4* __restrict__ src, float4* __restrict__ dst, uint32_t width, uint32_t height,
                               uint32_t block, uint32_t nranks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)width * height) return;
    const uint32_t y = (uint32_t)(i / width), x = (uint32_t)(i % width);
    const uint32_t b = y / block, rank = b % nranks, l = (b / nranks) * block + y % block;
    dst[i] = src[((size_t)rank * (height / nranks) + l) * width + x];
}

// ---------------------------------------------------------------- synthetic soup
// value j of triangle i = rand-conversion of pcg(seed + (16 i + j) * 0x9E3779B9); one mul and one
// add per value, no FMA (this file is built with --fmad=false) -> bit-identical to the host definition the tests hold.
__global__ void k_soup(uint32_t ntris, uint32_t seed, float scale, float* __restrict__ verts,
                       uint32_t* __restrict__ idx, float* __restrict__ faces) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntris) return;
    float f[15];
#pragma unroll
    for (uint32_t j = 0; j < 15; ++j) {
        uint32_t st = seed + (16u * i + j) * 0x9E3779B9u;
        f[j] = __uint2float_rn(bpt_pcg(st)) * 2.3283064365386963e-10f;
    }
    const float c[3] = {f[0] * 2.0f - 1.0f, f[1] * 2.0f - 2.0f, f[2] * 2.0f - 1.0f};
#pragma unroll
    for (int v = 0; v < 3; ++v)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float off = (f[3 + 3 * v + a] * 2.0f - 1.0f) * scale;
            verts[9 * (size_t)i + 3 * v + a] = c[a] + off;
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) faces[6 * (size_t)i + a] = f[12 + a] * 0.8f + 0.1f;
    const bool em = (i % 128u) == 0u;
    faces[6 * (size_t)i + 3] = em ? 17.0f : 0.0f;
    faces[6 * (size_t)i + 4] = em ? 12.0f : 0.0f;
    faces[6 * (size_t)i + 5] = em ? 4.0f : 0.0f;
    idx[3 * (size_t)i] = 3 * i; idx[3 * (size_t)i + 1] = 3 * i + 1; idx[3 * (size_t)i + 2] = 3 * i + 2;
}

// ---------------------------------------------------------------- scene front-end (reference main.cpp:37-57)
__global__ void k_obj_arrays(const float* __restrict__ positions, uint32_t npositions, const int32_t* __restrict__ corner_vertex,
                             uint32_t ncorners, const int32_t* __restrict__ face_material, const float* __restrict__ materials,
                             uint32_t nmaterials, float* __restrict__ verts, uint32_t* __restrict__ idx,
                             float* __restrict__ faces, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncorners) return;
    const int32_t vi = corner_vertex[i];
    if (vi < 0 || (uint32_t)vi >= npositions) { atomicAdd(bad, 1u); return; }
    verts[3 * (size_t)i + 0] = positions[3 * (size_t)vi + 0];    // main.cpp:42
    verts[3 * (size_t)i + 1] = -positions[3 * (size_t)vi + 1];   // :43 (the Y flip)
    verts[3 * (size_t)i + 2] = positions[3 * (size_t)vi + 2];    // :44
    idx[i] = i;                                                  // :45
    if (i % 3u == 0u) {
        const uint32_t f = i / 3u;
        const int32_t m = face_material[f];
        if (m < 0 || (uint32_t)m >= nmaterials) { atomicAdd(bad, 1u); return; }
#pragma unroll
        for (int k = 0; k < 6; ++k) faces[6 * (size_t)f + k] = materials[6 * (size_t)m + k];   // :52-55
    }
}

__global__ void k_check_indices(const uint32_t* __restrict__ idx, uint32_t n, uint32_t nverts, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || idx[i] < nverts) return;
    if (atomicAdd(&out[0], 1u) == 0u) out[1] = i;  // any out-of-range position serves the message
}

inline unsigned grid_for(uint64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace

void launch_generate(const FrameParams& p, const int32_t* frame_dev, uint32_t s0, uint32_t ns, uint32_t path_base, PathQueue q,
                     uint32_t* counts, uint32_t* fetch, uint32_t ncounters, cudaStream_t st) {
    const uint64_t threads = std::max<uint64_t>((uint64_t)tile_local_rows(p) * p.width * ns, ncounters);  // the first threads also reset the counters
    k_generate<<<grid_for(threads), kBlock, 0, st>>>(p, frame_dev, s0, ns, path_base, q, counts, fetch, ncounters);
}
void launch_gather_pass(uint32_t npix, uint32_t ns, float4* path_color, float4* frame_sum, cudaStream_t st) {
    k_gather_pass<<<grid_for(npix), kBlock, 0, st>>>(npix, ns, path_color, frame_sum);
}
void launch_shade(const FrameParams& p, const SceneView& s, const NeeView& nv, uint32_t depth, PathQueue in, const uint4* hits,
                  PathQueue out, uint32_t* counts, uint32_t* fetch, float4* path_color, uint32_t max_paths, unsigned num_sms,
                  cudaStream_t st) {
    const unsigned full = grid_for(max_paths);
    if (p.nee || p.rr_start_depth)
        k_shade<true><<<std::min(full, num_sms * 16u), kBlock, 0, st>>>(p, s, depth, in, hits, out, counts, fetch + kCounterStride + depth,
                                                                      path_color, nv.pdf_prev, nv.light_area);
    else
        k_shade<false><<<std::min(full, num_sms * 16u), kBlock, 0, st>>>(p, s, depth, in, hits, out, counts, fetch + kCounterStride + depth,
                                                                       path_color, nv.pdf_prev, nv.light_area);
}

void launch_nee(const FrameParams& p, const SceneView& s, const NeeView& nv, uint32_t depth, PathQueue in, const uint4* hits,
                const uint32_t* counts, float4* shadow_rays, float4* shadow_contrib, unsigned long long* ray_stat,
                uint32_t max_paths, unsigned num_sms, cudaStream_t st) {
    k_nee<<<std::min(grid_for(max_paths), num_sms * 16u), kBlock, 0, st>>>(p, s, nv, depth, in, hits, counts, shadow_rays,
                                                                          shadow_contrib, ray_stat);
}
void launch_nee_resolve(uint32_t depth, const uint32_t* counts, const uint4* shadow_hits, const float4* shadow_contrib,
                        const uint32_t* pixel, float4* path_color, uint32_t max_paths, unsigned num_sms, cudaStream_t st) {
    k_nee_resolve<<<std::min(grid_for(max_paths), num_sms * 16u), kBlock, 0, st>>>(depth, counts, shadow_hits, shadow_contrib,
                                                                                  pixel, path_color);
}
static __global__ void k_set_i32(int32_t* dst, int32_t v) { *dst = v; }
void launch_set_i32(int32_t* dst, int32_t v, cudaStream_t st) { k_set_i32<<<1, 1, 0, st>>>(dst, v); }
void launch_shade_records(const float* verts, const uint32_t* idx, const float* faces, uint32_t ntris, float4* out, cudaStream_t st) {
    k_shade_records<<<grid_for(ntris), kBlock, 0, st>>>(verts, idx, faces, ntris, out);
}
void launch_refine_hits(const SceneView& s, const float4* rays, uint4* hits, uint32_t n, cudaStream_t st) {
    k_refine_hits<<<grid_for(n), kBlock, 0, st>>>(s, rays, hits, n);
}
void launch_accumulate(const FrameParams& p, const int32_t* frame_dev, float4* frame_sum, float4* image, cudaStream_t st) {
    k_accumulate<<<grid_for((uint64_t)tile_local_rows(p) * p.width), kBlock, 0, st>>>(p, frame_dev, frame_sum, image);
}
void launch_obj_arrays(const float* positions, uint32_t npositions, const int32_t* corner_vertex, uint32_t ncorners,
                       const int32_t* face_material, const float* materials, uint32_t nmaterials, float* verts, uint32_t* idx,
                       float* faces, uint32_t* bad, cudaStream_t st) {
    k_obj_arrays<<<grid_for(ncorners), kBlock, 0, st>>>(positions, npositions, corner_vertex, ncorners, face_material, materials,
                                                        nmaterials, verts, idx, faces, bad);
}
void launch_check_indices(const uint32_t* idx, uint32_t n, uint32_t nverts, uint32_t* out, cudaStream_t st) {
    k_check_indices<<<grid_for(n), kBlock, 0, st>>>(idx, n, nverts, out);
}
void launch_soup(uint32_t ntris, uint32_t seed, float scale, float* verts, uint32_t* idx, float* faces, cudaStream_t st) {
    k_soup<<<grid_for(ntris), kBlock, 0, st>>>(ntris, seed, scale, verts, idx, faces);
}
void launch_deinterleave(const float4* rank_major, float4* row_major, uint32_t width, uint32_t height, uint32_t block,
                         uint32_t nranks, cudaStream_t st) {
    k_deinterleave<<<grid_for((uint64_t)width * height), kBlock, 0, st>>>(rank_major, row_major, width, height, block, nranks);
}
void launch_image_to_bgra8(const float4* image, uint8_t* bgra, size_t npix, cudaStream_t st) {
    k_image_to_bgra8<<<grid_for(npix), kBlock, 0, st>>>(image, bgra, npix);
}
