// ref_shade_glue.cpp — builds oracle/_ref/libref_shade.so: the reference's OWN shader text, compiled as C++.
//
// *** TEST INFRASTRUCTURE ONLY *** — the pin of oracle/oracle.cpp for everything the reference's source states:
// shaders/common.glsl:13-37 (pcg, pcg2d, rand), raygen.rgen:14-39 (sampling) and :41-91 (sample loop, camera, path
// update, running mean), closesthit.rchit:24-65, miss.rmiss:8-12. The four files are read from the reference
// checkout where they lie, passed through oracle/glsl_to_cpp.py (rules R1-R7 there) into oracle/_ref/gen/*.inc and
// included below, one namespace per shader stage, under oracle/glsl_shim.h. What is NOT reference code here is the
// part the reference does not contain either: the driver's acceleration-structure traversal behind traceRayEXT
// (closed source). The glue binds it to an intersector: a brute-force float Moeller-Trumbore loop over the bound
// vertex / index buffers, or any callback with the same contract (closest opaque hit in [tmin, tmax], no culling,
// barycentrics of v1 and v2; raygen.rgen:63-75, main.cpp:525) — tests pass the oracle's own intersector so that the
// two renders differ only where the restatement of the shader text differs from the text.
//
// Build: `make -C oracle _ref` (needs /root/reference; the built .so travels to the GPU box, the sources do not).
#include <atomic>
#include <thread>
#include <vector>

#include "glsl_shim.h"

namespace glsl {
thread_local uvec3 gl_LaunchIDEXT, gl_LaunchSizeEXT;
thread_local int gl_PrimitiveID = 0;
int g_spp_override = 0, g_depth_override = 0;
}  // namespace glsl

#define main shader_main
namespace rgen {
GLSL_USING_BUILTINS
#include "_ref/gen/raygen.rgen.inc"
}  // namespace rgen
namespace rchit {
GLSL_USING_BUILTINS
#include "_ref/gen/closesthit.rchit.inc"
}  // namespace rchit
namespace rmiss {
GLSL_USING_BUILTINS
#include "_ref/gen/miss.rmiss.inc"
}  // namespace rmiss
#undef main

namespace {
thread_local uint64_t t_rays = 0;
uint32_t g_ntris = 0;  // triangles in the bound index buffer

// rayPayloadEXT (raygen) and rayPayloadInEXT (hit / miss) at location 0 are one object in the pipeline
template <class A, class B>
inline void copy_payload(A& dst, const B& src) {
    dst.position = src.position; dst.normal = src.normal; dst.emission = src.emission; dst.brdf = src.brdf;
    dst.done = src.done;
}

// default intersector: every triangle of the bound buffers, float Moeller-Trumbore, either side; equal distances
// resolve to the lowest primitive id (the asset holds exact duplicates)
void brute_force(void*, const float* o, float tmin, const float* d, float tmax, glsl::RefHit* out) {
    using glsl::vec3;
    const vec3 O(o[0], o[1], o[2]), D(d[0], d[1], d[2]);
    glsl::RefHit h{tmax, 0.f, 0.f, 0xffffffffu};
    for (uint32_t i = 0; i < g_ntris; ++i) {
        const float* a = rchit::vertices.data + 3 * size_t(rchit::indices.data[3 * size_t(i)]);
        const float* b = rchit::vertices.data + 3 * size_t(rchit::indices.data[3 * size_t(i) + 1]);
        const float* c = rchit::vertices.data + 3 * size_t(rchit::indices.data[3 * size_t(i) + 2]);
        const vec3 v0(a[0], a[1], a[2]);
        const vec3 e1 = vec3(b[0], b[1], b[2]) - v0, e2 = vec3(c[0], c[1], c[2]) - v0;
        const vec3 p = glsl::cross(D, e2);
        const float det = glsl::dot(e1, p);
        if (det == 0.0f) continue;
        const float inv = 1.0f / det;
        const vec3 s = O - v0;
        const float u = glsl::dot(s, p) * inv;
        if (!(u >= 0.0f) || u > 1.0f) continue;
        const vec3 q = glsl::cross(s, e1);
        const float v = glsl::dot(D, q) * inv;
        if (!(v >= 0.0f) || u + v > 1.0f) continue;
        const float t = glsl::dot(e2, q) * inv;
        if (!(t >= tmin) || !(t <= h.t)) continue;
        if (t < h.t || h.prim == 0xffffffffu) h = {t, u, v, i};
    }
    *out = h;
}
}  // namespace

// traceRayEXT (raygen.rgen:63-75): closest hit -> closest-hit stage with gl_PrimitiveID / attribs, else miss stage
void glsl::traceRayEXT(accelerationStructureEXT& as, uint, uint, uint, uint, uint, vec3 origin, float tmin, vec3 direction,
                       float tmax, int) {
    RefHit h;
    const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {direction.x, direction.y, direction.z};
    as.fn(as.user, o, tmin, d, tmax, &h);
    ++t_rays;
    if (h.prim != 0xffffffffu) {
        copy_payload(rchit::payload, rgen::payload);
        gl_PrimitiveID = int(h.prim);
        rchit::attribs = vec2(h.u, h.v);
        rchit::shader_main();
        copy_payload(rgen::payload, rchit::payload);
    } else {
        copy_payload(rmiss::payload, rgen::payload);
        rmiss::shader_main();
        copy_payload(rgen::payload, rmiss::payload);
    }
}

extern "C" {

// One vkCmdTraceRaysKHR(width, height, 1) (main.cpp:659) with push constant `frame` over the storage image `image`
// (width*height*4 floats, read-modify-write; rgba8 != 0: the reference's unorm8 image, else a float image).
// rows row0, row0 + row_step, ... < row1 only (row1 = 0: to the last row): every invocation is independent, so a subset
// of the rows is a subset of the same launch.
// spp_override / depth_override: 0 = the literals of the shader text (32, 8). intersect: NULL = built-in brute force.
// Returns the number of traceRayEXT calls.
uint64_t ref_shade_render(const float* verts, const uint32_t* indices, uint32_t nindices, const float* faces,
                          uint32_t width, uint32_t height, uint32_t row0, uint32_t row1, uint32_t row_step, int frame,
                          int spp_override,
                          int depth_override, int rgba8, glsl::ref_intersect_fn intersect, void* user, int nthreads,
                          float* image) {
    rchit::vertices.data = verts;
    rchit::indices.data = indices;
    rchit::faces.data = faces;
    g_ntris = nindices / 3;
    rgen::topLevelAS.fn = intersect ? intersect : brute_force;
    rgen::topLevelAS.user = user;
    rgen::outputImage.texels = image;
    rgen::outputImage.width = width;
    rgen::outputImage.height = height;
    rgen::outputImage.rgba8 = rgba8 != 0;
    rgen::frame = frame;
    glsl::g_spp_override = spp_override;
    glsl::g_depth_override = depth_override;
    if (row1 == 0 || row1 > height) row1 = height;
    if (row_step == 0) row_step = 1;
    if (nthreads <= 0) nthreads = int(std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<uint32_t> next{row0};
    std::atomic<uint64_t> total{0};
    auto worker = [&]() {
        t_rays = 0;
        glsl::gl_LaunchSizeEXT = glsl::uvec3(width, height, 1);
        for (;;) {
            const uint32_t y = next.fetch_add(row_step);
            if (y >= row1) break;
            for (uint32_t x = 0; x < width; ++x) {
                glsl::gl_LaunchIDEXT = glsl::uvec3(x, y, 0);
                rgen::shader_main();
            }
        }
        total += t_rays;
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return total.load();
}

// the integer RNG of common.glsl, as compiled from the text (KAT-1 cross-check)
uint32_t ref_pcg(uint32_t* state) { return rgen::pcg(*state); }
void ref_pcg2d(uint32_t* x, uint32_t* y) {
    glsl::uvec2 v = rgen::pcg2d(glsl::uvec2(*x, *y));
    *x = v.x; *y = v.y;
}
float ref_rand(uint32_t* seed) { return rgen::rand(*seed); }
// raygen.rgen:14-39
void ref_sample_direction(float r1, float r2, const float* n, float* out) {
    glsl::vec3 d = rgen::sampleDirection(r1, r2, glsl::vec3(n[0], n[1], n[2]));
    out[0] = d.x; out[1] = d.y; out[2] = d.z;
}

}  // extern "C"
