// vecmath.h -- the handful of fp32 vector operations the host mirror needs, written out with
// glm's generic (non-SIMD) operation order so that Camera / CubicSpline reproduce what the
// reference computes through glm (which is not vendored and not available here).
#pragma once

#include <cmath>

namespace vr {

struct ivec2 { int x = 0, y = 0; };
struct ivec3 { int x = 0, y = 0, z = 0;
               bool operator==(const ivec3& o) const { return x == o.x && y == o.y && z == o.z; } };
struct vec3  { float x = 0, y = 0, z = 0;
               bool operator==(const vec3& o) const { return x == o.x && y == o.y && z == o.z; } };

struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline vec4 operator+(const vec4& a, const vec4& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator-(const vec4& a, const vec4& b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline vec4 operator-(const vec4& a) { return {-a.x, -a.y, -a.z, -a.w}; }
inline vec4 operator*(const vec4& a, const vec4& b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline vec4 operator/(const vec4& a, const vec4& b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
inline vec4 operator*(float s, const vec4& a) { return {s * a.x, s * a.y, s * a.z, s * a.w}; }
inline vec4 operator*(const vec4& a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }

// glm::dot(vec4) pairs the terms: (x*x + y*y) + (z*z + w*w)
inline float dot4(const vec4& a, const vec4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
// glm::normalize(v) = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
inline vec4 normalize4(const vec4& v) { return v * (1.0f / std::sqrt(dot4(v, v))); }
// glm::length(vec3) = sqrt(x*x + y*y + z*z)
inline float length3(float x, float y, float z) { return std::sqrt(x * x + y * y + z * z); }
// glm::cross
inline vec4 cross3(const vec4& a, const vec4& b)
{
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y, 0.0f};
}

struct mat4 {
    vec4 col[4];   // column-major, like glm::mat4
    mat4() { col[0] = {1, 0, 0, 0}; col[1] = {0, 1, 0, 0}; col[2] = {0, 0, 1, 0}; col[3] = {0, 0, 0, 1}; }
    mat4(const vec4& a, const vec4& b, const vec4& c, const vec4& d) { col[0] = a; col[1] = b; col[2] = c; col[3] = d; }
};

constexpr float kPi = 3.14159265358979323846264338327950288f;   // glm::pi<float>()

}  // namespace vr
