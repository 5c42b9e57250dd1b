/*
 * oracle/shim/glm/glm.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * Minimal stand-in for the subset of GLM that /root/reference/src/Camera.cpp and
 * /root/reference/src/CubicSpline.cpp use, so that those two reference files compile UNMODIFIED where
 * they lie (oracle/Makefile, target _ref/libhost_ref.so).  GLM itself is a third-party dependency that
 * the reference neither vendors nor pins (no submodule, no CMake/conan manifest); what follows restates
 * GLM's published generic (non-SIMD, GLM_FORCE_PURE) definitions, 0.9.8 / 0.9.9 series, one IEEE binary32
 * operation per operator:
 *   dot(vec3)      = x*x + y*y + z*z                  (detail/func_geometric.inl, compute_dot<vec3>)
 *   dot(vec4)      = (x*x + y*y) + (z*z + w*w)        (compute_dot<vec4>)
 *   length(v)      = sqrt(dot(v,v))                   inversesqrt(x) = 1/sqrt(x)
 *   normalize(v)   = v * inversesqrt(dot(v,v))
 *   cross(x,y)     = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
 *   clamp(x,a,b)   = min(max(x,a),b);  min(x,y) = y<x ? y : x;  max(x,y) = x<y ? y : x
 *   mat4 * vec4    = (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)           (detail/type_mat4x4.inl)
 *   rotate(m,a,v)  = gtc/matrix_transform.inl (axis-angle, see below)
 * "parity unpinned" at this boundary: a different GLM build (SIMD paths) may differ in the last ulp.
 */
#pragma once

#include <cmath>
#include <type_traits>

namespace glm {

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    template <class A, class B, class C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> explicit vec3(S s) : x((float)s), y((float)s), z((float)s) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
inline vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }

struct ivec3 { int x, y, z; ivec3() : x(0), y(0), z(0) {} ivec3(int a, int b, int c) : x(a), y(b), z(c) {} };
inline bool operator==(const ivec3& a, const ivec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} };
struct uvec3 { unsigned x, y, z; uvec3() : x(0), y(0), z(0) {} uvec3(unsigned a, unsigned b, unsigned c) : x(a), y(b), z(c) {} };

struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    template <class A, class B, class C, class D> vec4(A a, B b, C c, D d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> explicit vec4(S s) : x((float)s), y((float)s), z((float)s), w((float)s) {}
    template <class S> vec4(const vec3& v, S w_) : x(v.x), y(v.y), z(v.z), w((float)w_) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator/(const vec4& a, const vec4& b) { return vec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }
inline vec4 operator*(float s, const vec4& v) { return vec4(s * v.x, s * v.y, s * v.z, s * v.w); }
inline vec4 operator*(const vec4& v, float s) { return vec4(v.x * s, v.y * s, v.z * s, v.w * s); }
inline vec4& operator+=(vec4& a, const vec4& b) { a = a + b; return a; }
inline vec4& operator-=(vec4& a, const vec4& b) { a = a - b; return a; }

struct mat4 {
    vec4 c[4];                                            /* columns */
    mat4() { c[0] = vec4(0, 0, 0, 0); c[1] = c[0]; c[2] = c[0]; c[3] = c[0]; }
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>>
    explicit mat4(S s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    mat4(const vec4& a, const vec4& b, const vec4& d, const vec4& e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    const vec4 Mul0 = m[0] * vec4(v.x), Mul1 = m[1] * vec4(v.y);
    const vec4 Add0 = Mul0 + Mul1;
    const vec4 Mul2 = m[2] * vec4(v.z), Mul3 = m[3] * vec4(v.w);
    const vec4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

inline float min(float x, float y) { return (y < x) ? y : x; }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

inline float dot(const vec3& a, const vec3& b) { const vec3 t(a.x * b.x, a.y * b.y, a.z * b.z); return t.x + t.y + t.z; }
inline float dot(const vec4& a, const vec4& b) { const vec4 t(a * b); return (t.x + t.y) + (t.z + t.w); }
inline float length(const vec3& v) { return sqrt(dot(v, v)); }
inline float length(const vec4& v) { return sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec4 normalize(const vec4& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 cross(const vec3& x, const vec3& y)
{
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}

}  // namespace glm
