// glsl_shim.h — the minimum of GLSL 4.60 + GL_EXT_ray_tracing needed to compile the reference's shader TEXT as C++.
//
// *** TEST INFRASTRUCTURE ONLY *** (see oracle/oracle.cpp). Used by oracle/ref_shade_glue.cpp, which #includes the
// reference's shaders/common.glsl, raygen.rgen, closesthit.rchit and miss.rmiss — read from /root/reference where
// they lie, passed through the mechanical token transform oracle/glsl_to_cpp.py (qualifiers, float literal suffixes,
// interface blocks, left-to-right argument evaluation) and written to oracle/_ref/gen/ (git-ignored) — into one
// namespace per shader stage and builds oracle/_ref/libref_shade.so from them. Nothing of the reference is copied
// into this repository: this header only supplies the language the shader text is written in.
//
// Semantics the shim has to choose (GLSL leaves them to the implementation; the CUDA path and oracle.cpp choose the same):
//   * float is IEEE binary32, one rounding per operation, no contraction (built with -ffp-contract=off)
//   * dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z, evaluated left to right
//   * normalize(v) = v / sqrt(dot(v,v)); cross(a,b) componentwise (a.y*b.z - a.z*b.y, ...)
//   * sin / cos / sqrt = the host libm's float versions
//   * ivec2 * uint converts the ivec2 to uvec2 first (GLSL 4.60 implicit conversions, section 4.1.10); the shipped
//     raygen.rgen.spv does the same (an OpIMul on u32)
//   * function-call and constructor arguments are evaluated left to right (GLSL 4.60 section 6.1.1 "in order, from
//     left to right"): C++ leaves that order unspecified, so the transform routes every call with two or more
//     arguments through GLSL_CALL, which evaluates them inside a braced initialiser list ([dcl.init.list]/4)
//   * an rgba8 image (raygen.rgen:7) stores unorm8: clamp to [0,1], * 255, round to nearest even; a float image mode
//     exists beside it because the north star asks for a float4 accumulation
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <tuple>
#include <utility>

#ifdef M_PI
#undef M_PI  // common.glsl:11 declares its own `const highp float M_PI`
#endif

namespace glsl {

typedef unsigned int uint;

struct vec2; struct vec3; struct vec4; struct uvec2; struct uvec3; struct ivec2;

// swizzle proxy: aliases the first N components of its parent (read: conversion to V, write: assignment from V)
template <class V, class T, int N>
struct swz {
    T e[N];
    operator V() const;
    swz& operator=(const V& v);
};

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(float a) : x(a), y(a) {}
    explicit vec2(const uvec2& u);
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
};
struct vec4 {
    union {
        struct { float x, y, z, w; };
        swz<vec3, float, 3> xyz;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
};
struct uvec2 {
    uint x, y;
    uvec2() : x(0), y(0) {}
    uvec2(uint a, uint b) : x(a), y(b) {}
    explicit uvec2(const ivec2& v);
};
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const uvec2& u) : x(int(u.x)), y(int(u.y)) {}
};
struct uvec3 {
    union {
        struct { uint x, y, z; };
        swz<uvec2, uint, 2> xy;
    };
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
    uvec3(const uvec3& o) : x(o.x), y(o.y), z(o.z) {}
    uvec3& operator=(const uvec3& o) { x = o.x; y = o.y; z = o.z; return *this; }
};
inline vec2::vec2(const uvec2& u) : x(float(u.x)), y(float(u.y)) {}
inline uvec2::uvec2(const ivec2& v) : x(uint(v.x)), y(uint(v.y)) {}
template <> inline swz<vec3, float, 3>::operator vec3() const { return vec3(e[0], e[1], e[2]); }
template <> inline swz<vec3, float, 3>& swz<vec3, float, 3>::operator=(const vec3& v) { e[0] = v.x; e[1] = v.y; e[2] = v.z; return *this; }
template <> inline swz<uvec2, uint, 2>::operator uvec2() const { return uvec2(e[0], e[1]); }
template <> inline swz<uvec2, uint, 2>& swz<uvec2, uint, 2>::operator=(const uvec2& v) { e[0] = v.x; e[1] = v.y; return *this; }

// ---- operators (componentwise; a scalar operand is applied to every component)
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator-(const vec2& a, float s) { return vec2(a.x - s, a.y - s); }

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, const vec3& b) { a = a * b; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }

inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

inline uvec2 operator*(const uvec2& a, uint s) { return uvec2(a.x * s, a.y * s); }
inline uvec2 operator+(const uvec2& a, uint s) { return uvec2(a.x + s, a.y + s); }
inline uvec2 operator>>(const uvec2& a, uint s) { return uvec2(a.x >> s, a.y >> s); }
inline uvec2 operator^(const uvec2& a, const uvec2& b) { return uvec2(a.x ^ b.x, a.y ^ b.y); }
inline uvec2 operator*(const ivec2& a, uint s) { return uvec2(a) * s; }  // implicit ivec2 -> uvec2 (header comment)

// ---- built-in functions
inline float sqrt(float x) { return std::sqrt(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float abs(float x) { return std::fabs(x); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline vec3 normalize(const vec3& a) { return a / sqrt(dot(a, a)); }

// ---- left-to-right argument evaluation
template <class... T>
struct Args {
    std::tuple<T&&...> t;
    Args(T&&... a) : t(std::forward<T>(a)...) {}
};
template <class... T> Args(T&&...) -> Args<T...>;
template <class F, class... T>
inline decltype(auto) invoke_ltr(F&& f, Args<T...>&& a) { return std::apply(std::forward<F>(f), std::move(a.t)); }
#define GLSL_CALL(F, ...) \
    ::glsl::invoke_ltr([&](auto&&... glsl_a) -> decltype(auto) { return F(std::forward<decltype(glsl_a)>(glsl_a)...); }, ::glsl::Args{__VA_ARGS__})

// ---- storage: SSBO unsized arrays, the storage image, the acceleration structure
template <class T>
struct buffer_array {
    const T* data = nullptr;
    const T& operator[](uint i) const { return data[i]; }
};

struct image2D {
    float* texels = nullptr;  // width * height * 4 floats; in rgba8 mode every stored value is k/255
    uint width = 0, height = 0;
    bool rgba8 = true;        // the reference's format (raygen.rgen:7, main.cpp:481-484)
};
inline float unorm8(float v) {
    float c = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    if (v != v) c = 0.0f;
    return std::nearbyint(c * 255.0f) / 255.0f;
}
inline vec4 imageLoad(const image2D& img, const ivec2& p) {
    const float* t = img.texels + 4 * (size_t(p.y) * img.width + size_t(p.x));
    return vec4(t[0], t[1], t[2], t[3]);
}
inline void imageStore(image2D& img, const ivec2& p, const vec4& v) {
    float* t = img.texels + 4 * (size_t(p.y) * img.width + size_t(p.x));
    if (img.rgba8) { t[0] = unorm8(v.x); t[1] = unorm8(v.y); t[2] = unorm8(v.z); t[3] = unorm8(v.w); }
    else { t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w; }
}

// closest-hit query the glue binds to an intersector (the driver's traversal is closed source: SURVEY.md 8c)
struct RefHit { float t, u, v; uint32_t prim; };
typedef void (*ref_intersect_fn)(void* user, const float* origin, float tmin, const float* dir, float tmax, RefHit* out);
struct accelerationStructureEXT {
    ref_intersect_fn fn = nullptr;
    void* user = nullptr;
};
constexpr uint gl_RayFlagsOpaqueEXT = 1u;
void traceRayEXT(accelerationStructureEXT& as, uint rayFlags, uint cullMask, uint sbtRecordOffset, uint sbtRecordStride,
                 uint missIndex, vec3 origin, float tmin, vec3 direction, float tmax, int payloadLocation);

// per-invocation built-ins (one shader invocation per host thread at a time)
extern thread_local uvec3 gl_LaunchIDEXT, gl_LaunchSizeEXT;
extern thread_local int gl_PrimitiveID;

// The two literals the transform keeps overridable so that BASELINE.json's "1 spp, depth 2" plumbing configuration
// can be rendered from the same text: `int maxSamples = 32;` (raygen.rgen:43) and `depth < 8` (raygen.rgen:62)
// become ref_spp(32) / ref_depth(8), which return their argument unless an override is set.
extern int g_spp_override, g_depth_override;
inline int ref_spp(int text_value) { return g_spp_override > 0 ? g_spp_override : text_value; }
inline uint ref_depth(uint text_value) { return g_depth_override > 0 ? uint(g_depth_override) : text_value; }

}  // namespace glsl

// qualifiers that have no meaning on the host
#define highp
#define layout(...)
#define uniform
#define rayPayloadEXT thread_local
#define rayPayloadInEXT thread_local
#define hitAttributeEXT thread_local

// names a shader stage sees unqualified (using-declarations: they hide ::abs(int) and friends)
#define GLSL_USING_BUILTINS                                                                                       \
    using glsl::uint; using glsl::vec2; using glsl::vec3; using glsl::vec4; using glsl::uvec2; using glsl::uvec3;  \
    using glsl::ivec2; using glsl::sqrt; using glsl::sin; using glsl::cos; using glsl::abs; using glsl::dot;       \
    using glsl::cross; using glsl::normalize; using glsl::image2D; using glsl::imageLoad; using glsl::imageStore;  \
    using glsl::accelerationStructureEXT; using glsl::traceRayEXT; using glsl::gl_RayFlagsOpaqueEXT;               \
    using glsl::gl_LaunchIDEXT; using glsl::gl_LaunchSizeEXT; using glsl::gl_PrimitiveID;
