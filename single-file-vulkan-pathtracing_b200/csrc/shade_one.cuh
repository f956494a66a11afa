// shade_one.cuh — the per-path device code shared by the wavefront shade kernel (shade.cu) and the fused path kernel
// (trace.cu): primary-ray generation (shaders/raygen.rgen:47-60) and one shade step (closesthit.rchit:50-64,
// miss.rmiss:8-12, raygen.rgen:14-39,76-83).
//
// Every translation unit that includes this header MUST be compiled with --fmad=false and without fast-math: each
// float operation below is the single IEEE operation the shader text names, in the order the CPU checker (tests)
// restates it; only sinf/cosf may differ from a host libm in the last bit.
#pragma once
#include "shade.cuh"

namespace bpt_shade {

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ V3 normalize(V3 a) { return a / sqrtf(dot(a, a)); }

constexpr float kTwoPi = 6.2831855f, kPi = 3.1415927f, kPdf = 0.15915494f;

// raygen.rgen:47-57 for path i of a sample pass: the pass carries samples s0.. of every tile pixel, path i = sample
// slot i / npix of tile-local pixel i % npix. Returns the ray and the seed after the two jitter draws.
__device__ __forceinline__ void gen_primary(const FrameParams& p, uint32_t s0, uint32_t npix, uint32_t i, float4& ro, float4& rd,
                                            uint32_t& seed_out) {
    const uint32_t slot = i / npix, pl = i - slot * npix;
    const uint32_t px = pl % p.width, py = tile_global_row(p, pl / p.width);
    const uint32_t k = (s0 + slot) + p.spp_per_frame * (uint32_t)p.frame + 1u;  // raygen.rgen:47
    uint32_t sx = px * k, sy = py * k;
    bpt_pcg2d(sx, sy);
    uint32_t seed = sx + sy;
    const float r1 = bpt_rand(seed);
    const float r2 = bpt_rand(seed);
    const float scx = (float)px + r1, scy = (float)py + r2;            // :51
    const float ux = scx / (float)p.width, uy = scy / (float)p.height; // :52
    const float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;          // :53
    const V3 o{p.cam_origin[0], p.cam_origin[1], p.cam_origin[2]};
    const V3 target{dx + p.cam_target[0], dy + p.cam_target[1], p.cam_target[2]};
    const V3 d = normalize(target - o);                                // :57
    ro = make_float4(o.x, o.y, o.z, p.tmin);
    rd = make_float4(d.x, d.y, d.z, p.tmax);
    seed_out = seed;
}

__device__ __forceinline__ V3 xform(const float* m, V3 p) {
    if (!m) return p;
    return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
            m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}

// Barycentrics (u,v) of v1,v2 and the distance t at the ray/triangle intersection, Moeller-Trumbore with one IEEE
// operation per step (attribs of closesthit.rchit:56). The traversal kernel only decides WHICH triangle is closest
// (Woop form, fast, but it loses bits on slivers and far origins) and reports its own t; on a degenerate determinant
// u = v = 0 and that t are kept.
__device__ __forceinline__ void barycentrics(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float& u, float& v, float& t) {
    const V3 e1 = v1 - v0, e2 = v2 - v0;
    const V3 p = cross(d, e2);
    const float det = dot(e1, p);
    u = 0.0f; v = 0.0f;
    if (det == 0.0f) return;
    const float inv = 1.0f / det;
    const V3 s = o - v0;
    const V3 q = cross(s, e1);
    const float uu = dot(s, p) * inv, vv = dot(d, q) * inv, tt = dot(e2, q) * inv;
    if (isfinite(uu) && isfinite(vv) && isfinite(tt)) { u = uu; v = vv; t = tt; }
}

// path_color[path id] collects `color` of one sample (raygen.rgen:76); k_gather_pass folds the samples of a pass into
// the frame sum in sample order, so the result does not depend on how many samples a pass carries. One path is owned
// by one thread at a time: a plain load - add - store. (A vector reduction, RED.ADD.F32x4, would spare the thread the
// wait for the load, but the L2 atomic units sustain only ~7 G of them per second: measured +20 ms per 440 M-ray frame,
// Cornell box 12.1 -> 7.9 Gray/s.)
template <bool RED = false>
__device__ __forceinline__ void add_color(float4* path_color, uint32_t pix, V3 c) {
    if (RED) {
        // the fused path kernel: nothing else runs in the warp while it shades, so it must not wait for the load; one
        // vector reduction per contribution (a path contributes about once, far below the L2 atomic rate). A path's
        // contributions arrive one after the other, in bounce order, so the sum has the same value as below.
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(path_color + pix), "f"(c.x), "f"(c.y), "f"(c.z), "f"(0.0f) : "memory");
        return;
    }
    float4 acc = path_color[pix];
    acc.x += c.x; acc.y += c.y; acc.z += c.z;
    path_color[pix] = acc;
}

// What every shade kernel hands to shade_one for a path: its hit, state, path id, ray and the (object-space) shading
// record of the primitive it hit.
struct ShadeOut { float4 ro, rd, st; };
// closesthit.rchit:50-65 / miss.rmiss:8-12 / raygen.rgen:76-83 for one path. Returns true when the path continues
// (o holds its next ray and state). EXTRA: the instance that also knows the estimators the reference does not have
// (next-event estimation, Russian roulette); the reference's estimator runs the instance without that code (the extra
// branches and registers cost the Cornell box 6 % of its step).
template <bool EXTRA, bool RED = false>
__device__ __forceinline__ bool shade_one(const FrameParams& p, const SceneView& s, uint32_t depth, uint4 h, float4 st, uint32_t pix,
                                          float4 ro, float4 rd, float4 ra, float4 rb, float4 rc, float4 rdd, float4* path_color,
                                          float* pdf_prev, float light_area, ShadeOut& o) {
    V3 w{st.x, st.y, st.z};
    uint32_t seed = __float_as_uint(st.w);
    if (h.w == BPT_MISS) {
        // miss.rmiss:10-11 then raygen.rgen:76 and the break at :81
        add_color<RED>(path_color, pix, w * V3{p.sky[0], p.sky[1], p.sky[2]});
        return false;
    }
    const float* m = s.xforms ? s.xforms + 12 * (size_t)(h.w / s.ntris) : nullptr;
    const V3 v0 = xform(m, V3{ra.x, ra.y, ra.z}), v1 = xform(m, V3{ra.w, rb.x, rb.y}), v2 = xform(m, V3{rb.z, rb.w, rc.x});
    const V3 kd{rc.y, rc.z, rc.w}, ke{rdd.x, rdd.y, rdd.z};
    // the barycentrics that shading consumes are derived from the original vertices, so the hit position
    // carries no traversal-format error (the traversal kernel only names the closest triangle)
    float u, v, t = __uint_as_float(h.x);
    barycentrics(V3{ro.x, ro.y, ro.z}, V3{rd.x, rd.y, rd.z}, v0, v1, v2, u, v, t);
    const float b0 = 1.0f - u - v;                           // closesthit.rchit:56
    const V3 pos = v0 * b0 + v1 * u + v2 * v;                // :57
    const V3 nrm = -normalize(cross(v1 - v0, v2 - v0));      // :58, :43-48
    if (ke.x != 0.0f || ke.y != 0.0f || ke.z != 0.0f) {      // raygen.rgen:76 (adding 0 is exact)
        V3 c = w * ke;
        if (EXTRA && p.nee && depth > 0u) {
            // next-event estimation: the previous vertex sampled this emitter by area as well; balance heuristic
            // between the pdf the bounce direction was drawn with and the area sampler's pdf for this point
            const float cy = fabsf(dot(V3{rd.x, rd.y, rd.z}, nrm));
            const float pl = cy > 0.0f && light_area > 0.0f ? t * t / (cy * light_area) : 0.0f;
            const float pp = pdf_prev[pix];
            c = c * (pp / (pp + pl));
        }
        add_color<RED>(path_color, pix, c);
    }
    if (depth + 1u >= p.max_depth) return false;             // the next segment would not be traced
    const V3 brdf = kd / kPi;                                // closesthit.rchit:61
    const float r1 = bpt_rand(seed);
    const float r2 = bpt_rand(seed);                         // raygen.rgen:78
    V3 T, B;                                                 // :14-21
    if (fabsf(nrm.x) > fabsf(nrm.y)) T = V3{nrm.z, 0.0f, -nrm.x} / sqrtf(nrm.x * nrm.x + nrm.z * nrm.z);
    else T = V3{0.0f, -nrm.z, nrm.y} / sqrtf(nrm.y * nrm.y + nrm.z * nrm.z);
    B = cross(nrm, T);
    V3 l;
    if (p.sampler == BPT_SAMPLER_COSINE) {
        const float sr = sqrtf(r1);
        l = V3{cosf(kTwoPi * r2) * sr, sinf(kTwoPi * r2) * sr, sqrtf(1.0f - r1)};
    } else {                                                 // :23-30 uniform hemisphere
        const float sr = sqrtf(1.0f - r1 * r1);
        l = V3{cosf(kTwoPi * r2) * sr, sinf(kTwoPi * r2) * sr, r1};
    }
    const V3 d = l.x * T + l.y * B + l.z * nrm;              // :38
    if (EXTRA && p.nee) pdf_prev[pix] = p.sampler == BPT_SAMPLER_COSINE ? dot(d, nrm) / kPi : kPdf;
    if (p.sampler == BPT_SAMPLER_COSINE) w = w * (brdf * kPi);
    else w = w * (brdf * dot(d, nrm) / kPdf);                // :79-80
    if (EXTRA && p.rr_start_depth && depth + 1u >= p.rr_start_depth) {   // Russian roulette (bpt.h), not the reference
        const float q = fminf(1.0f, fmaxf(w.x, fmaxf(w.y, w.z)));
        const float r3 = bpt_rand(seed);
        if (!(r3 < q)) return false;
        w = w / q;
    }
    o.ro = make_float4(pos.x, pos.y, pos.z, p.tmin);
    o.rd = make_float4(d.x, d.y, d.z, p.tmax);
    o.st = make_float4(w.x, w.y, w.z, __uint_as_float(seed));
    return true;
}

}  // namespace bpt_shade
