/*
 * oracle/shim/glsl_compat.h -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).
 *
 * The part of GLSL 4.30 that /root/reference/VolumeRenderer.cs uses, as C++20, so that the
 * reference's OWN shader text can be compiled by g++ where it lies and run on the CPU
 * (oracle/Makefile, target _ref/libshader_ref.so).  Nothing of the shader is restated here:
 * this header only supplies the language -- vector types with the swizzles the shader spells
 * (.xyz .xzy .xy .rgb .a .r), the arithmetic operators with GLSL's int -> float promotion, and
 * the built-ins it calls.  The build adds -fsingle-precision-constant (a GLSL literal `0.95`
 * is a float, not a double) and -ffp-contract=off.
 *
 * What GLSL leaves to the implementation, and how it is defined here (the same definitions
 * as oracle/march_oracle.c; a real GPU driver may differ in the last ulp of these):
 *   length(v)    = sqrt(x*x + y*y + z*z [+ w*w])      summed left to right, IEEE sqrt
 *   normalize(v) = v / length(v)                      IEEE division per component
 *   mat4 * vec4  = ((c0*v.x + c1*v.y) + c2*v.z) + c3*v.w
 *   a / b        = IEEE division (GLSL allows 2.5 ulp)
 *   min(x,y) = y < x ? y : x ;  max(x,y) = x < y ? y : x ;  clamp = min(max(x,lo),hi)   (spec 8.3)
 *   texture(usampler3D, tc): GL does not filter integer textures (RendererCore.cpp:414-419 sets
 *     GL_LINEAR on R8UI/R16UI), so the sampler carries the filter explicitly:
 *     NEAREST   i = clamp(floor(u*N), 0, N-1)                       (de-facto driver behaviour)
 *     TRILINEAR f = fma(u,N,-0.5); i0 = floor(f); w = f - i0; texels i0, i0+1 clamped to the edge
 *               (GL_CLAMP_TO_EDGE, RendererCore.cpp:411-413); lerp(a,b,w) = fma(w, b-a, a) in
 *               x, then y, then z.  `.r` of the result is a float so that the
 *               filtered value survives `vec4(texture(...).r)`; for NEAREST it is the integer.
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <type_traits>

#define GLSL_ARITH(S) class = std::enable_if_t<std::is_arithmetic_v<S>>

struct vec3;
struct ivec3;
struct uvec2;

/* ---- swizzle proxies: live in a union with the vector's components ---- */
template <class V, class T, int A, int B, int C>
struct swz3 {
    T v[4];
    operator V() const { return V(v[A], v[B], v[C]); }
    swz3& operator*=(float s) { v[A] *= s; v[B] *= s; v[C] *= s; return *this; }
    swz3& operator=(const V& o) { const V t = o; v[A] = t.x; v[B] = t.y; v[C] = t.z; return *this; }
};
template <class V, class T, int A, int B>
struct swz2 {
    T v[4];
    operator V() const { return V(v[A], v[B]); }
};

struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz3<vec3, float, 0, 1, 2> xyz;
        swz3<vec3, float, 0, 2, 1> xzy;
    };
    vec3() : x(0), y(0), z(0) {}
    template <class S, GLSL_ARITH(S)> explicit vec3(S s) : x((float)s), y((float)s), z((float)s) {}
    template <class X, class Y, class Z> vec3(X x_, Y y_, Z z_) : x((float)x_), y((float)y_), z((float)z_) {}
};

struct ivec3 {
    union {
        struct { int x, y, z; };
        swz3<ivec3, int, 0, 1, 2> xyz;
        swz3<ivec3, int, 0, 2, 1> xzy;
    };
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
};

struct uvec2 {
    unsigned x, y;
    uvec2(unsigned x_, unsigned y_) : x(x_), y(y_) {}
};
struct uvec3 {
    union {
        struct { unsigned x, y, z; };
        swz2<uvec2, unsigned, 0, 1> xy;
    };
    uvec3() : x(0), y(0), z(0) {}
    uvec3(unsigned x_, unsigned y_, unsigned z_) : x(x_), y(y_), z(z_) {}
};
struct uvec4 {
    unsigned x, y, z, w;
    template <class S, GLSL_ARITH(S)> explicit uvec4(S s) : x((unsigned)s), y((unsigned)s), z((unsigned)s), w((unsigned)s) {}
};
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int x_, int y_) : x(x_), y(y_) {}
    explicit ivec2(const uvec2& u) : x((int)u.x), y((int)u.y) {}
};
struct bvec3 { bool x, y, z; };

/* anything that converts to a vec3 of floats / of ints */
template <class T> struct is_v3 : std::false_type {};
template <> struct is_v3<vec3> : std::true_type {};
template <int A, int B, int C> struct is_v3<swz3<vec3, float, A, B, C>> : std::true_type {};
template <class T> struct is_iv3 : std::false_type {};
template <> struct is_iv3<ivec3> : std::true_type {};
template <int A, int B, int C> struct is_iv3<swz3<ivec3, int, A, B, C>> : std::true_type {};

template <class T, std::enable_if_t<is_v3<T>::value, int> = 0> inline vec3 to_vec3(const T& t) { return (vec3)t; }
template <class T, std::enable_if_t<is_iv3<T>::value, int> = 0> inline vec3 to_vec3(const T& t)
{
    const ivec3 i = (ivec3)t;
    return vec3((float)i.x, (float)i.y, (float)i.z);      /* GLSL implicit int -> float */
}

struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz3<vec3, float, 0, 1, 2> xyz;
        swz3<vec3, float, 0, 2, 1> xzy;
        swz3<vec3, float, 0, 1, 2> rgb;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    template <class S, GLSL_ARITH(S)> explicit vec4(S s) : x((float)s), y((float)s), z((float)s), w((float)s) {}
    template <class X, class Y, class Z, class W_, GLSL_ARITH(X), GLSL_ARITH(W_)>
    vec4(X x_, Y y_, Z z_, W_ w_) : x((float)x_), y((float)y_), z((float)z_), w((float)w_) {}
    template <class V3, class S, std::enable_if_t<is_v3<V3>::value || is_iv3<V3>::value, int> = 0, GLSL_ARITH(S)>
    vec4(const V3& v, S w_) { const vec3 t = to_vec3(v); x = t.x; y = t.y; z = t.z; w = (float)w_; }
};

/* ---- vec4 arithmetic: one IEEE binary32 operation per component ---- */
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator/(const vec4& a, const vec4& b) { return vec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
template <class S, GLSL_ARITH(S)> inline vec4 operator*(const vec4& a, S s) { return a * vec4((float)s); }
template <class S, GLSL_ARITH(S)> inline vec4 operator*(S s, const vec4& a) { return vec4((float)s) * a; }
template <class S, GLSL_ARITH(S)> inline vec4 operator/(const vec4& a, S s) { return a / vec4((float)s); }
template <class S, GLSL_ARITH(S)> inline vec4 operator-(const vec4& a, S s) { return a - vec4((float)s); }
template <class S, GLSL_ARITH(S)> inline vec4 operator+(const vec4& a, S s) { return a + vec4((float)s); }
inline vec4& operator+=(vec4& a, const vec4& b) { a = a + b; return a; }
inline vec4& operator-=(vec4& a, const vec4& b) { a = a - b; return a; }
inline vec4& operator*=(vec4& a, const vec4& b) { a = a * b; return a; }
inline vec4& operator/=(vec4& a, const vec4& b) { a = a / b; return a; }
template <class S, GLSL_ARITH(S)> inline vec4& operator*=(vec4& a, S s) { a = a * s; return a; }
template <class S, GLSL_ARITH(S)> inline vec4& operator/=(vec4& a, S s) { a = a / s; return a; }

/* ---- vec3 arithmetic (operands may be swizzles) ---- */
#define GLSL_V3(L) std::enable_if_t<is_v3<L>::value, int> = 0
template <class L, class R, GLSL_V3(L), GLSL_V3(R)> inline vec3 operator+(const L& a_, const R& b_) { const vec3 a = a_, b = b_; return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class L, class R, GLSL_V3(L), GLSL_V3(R)> inline vec3 operator-(const L& a_, const R& b_) { const vec3 a = a_, b = b_; return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class L, class R, GLSL_V3(L), GLSL_V3(R)> inline vec3 operator*(const L& a_, const R& b_) { const vec3 a = a_, b = b_; return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class L, class R, GLSL_V3(L), GLSL_V3(R)> inline vec3 operator/(const L& a_, const R& b_) { const vec3 a = a_, b = b_; return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
template <class L, class S, GLSL_V3(L), GLSL_ARITH(S)> inline vec3 operator/(const L& a, S s) { return (vec3)a / vec3((float)s); }
template <class L, class S, GLSL_V3(L), GLSL_ARITH(S)> inline vec3 operator*(const L& a, S s) { return (vec3)a * vec3((float)s); }
template <class S, class R, GLSL_ARITH(S), GLSL_V3(R)> inline vec3 operator/(S s, const R& b) { return vec3((float)s) / (vec3)b; }
template <class S, class R, GLSL_ARITH(S), GLSL_V3(R)> inline vec3 operator*(S s, const R& b) { return vec3((float)s) * (vec3)b; }

/* ---- built-ins (spec 8.3, 8.5, 8.7) ---- */
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline int   min(int x, int y) { return (y < x) ? y : x; }
inline int   max(int x, int y) { return (x < y) ? y : x; }
inline float max(float x, int y) { return max(x, (float)y); }
inline float min(float x, int y) { return min(x, (float)y); }
inline vec4 min(const vec4& a, const vec4& b) { return vec4(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
inline vec4 max(const vec4& a, const vec4& b) { return vec4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
inline vec4 clamp(const vec4& x, const vec4& lo, const vec4& hi) { return min(max(x, lo), hi); }

template <class T, std::enable_if_t<is_v3<T>::value || is_iv3<T>::value, int> = 0>
inline float length(const T& t)
{
    const vec3 v = to_vec3(t);
    return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
}
inline float length(const vec4& v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w); }
inline vec4 normalize(const vec4& v) { return v / length(v); }

inline bvec3 greaterThan(const vec3& a, const vec3& b) { return bvec3{a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }

struct mat4 {
    vec4 c[4];      /* columns */
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w;
}

/* ---- image2D (rgba32f, write only) ---- */
struct image2D {
    float* data = nullptr;       /* W*H*4, row 0 = y 0 (GL image origin: bottom) */
    int w = 0, h = 0;
};
inline ivec2 imageSize(const image2D& im) { return ivec2(im.w, im.h); }
inline void imageStore(const image2D& im, const ivec2& p, const vec4& c)
{
    float* o = im.data + ((size_t)p.y * (size_t)im.w + (size_t)p.x) * 4;
    o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
}

/* ---- usampler3D ---- */
enum { GLSL_FILTER_NEAREST = 0, GLSL_FILTER_TRILINEAR = 1 };
struct usampler3D {
    const uint8_t* v8 = nullptr;
    const uint16_t* v16 = nullptr;
    int nx = 0, ny = 0, nz = 0;
    int filter = GLSL_FILTER_NEAREST;
    uint64_t* fetches = nullptr;          /* counts texture() calls */
    float texel(int x, int y, int z) const
    {
        const uint64_t i = ((uint64_t)z * (uint64_t)ny + (uint64_t)y) * (uint64_t)nx + (uint64_t)x;
        return v8 ? (float)v8[i] : (float)v16[i];
    }
};
struct utexel { float r, g, b, a; };
inline ivec3 textureSize(const usampler3D& s, int) { return ivec3(s.nx, s.ny, s.nz); }

inline int glsl_floor_to_int(float f)     /* GPU F2I semantics: NaN -> 0, saturating */
{
    if (!(f == f)) return 0;
    const float fl = std::floor(f);
    if (fl <= -2147483648.0f) return (int)(-2147483647 - 1);
    if (fl >= 2147483648.0f) return 2147483647;
    return (int)fl;
}
inline int glsl_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

inline utexel texture(const usampler3D& s, const vec3& tc)
{
    if (s.fetches) ++*s.fetches;
    float r;
    if (s.filter == GLSL_FILTER_NEAREST) {
        const int ix = glsl_clampi(glsl_floor_to_int(tc.x * (float)s.nx), 0, s.nx - 1);
        const int iy = glsl_clampi(glsl_floor_to_int(tc.y * (float)s.ny), 0, s.ny - 1);
        const int iz = glsl_clampi(glsl_floor_to_int(tc.z * (float)s.nz), 0, s.nz - 1);
        r = s.texel(ix, iy, iz);
    } else {
        const float fx = std::fmaf(tc.x, (float)s.nx, -0.5f);
        const float fy = std::fmaf(tc.y, (float)s.ny, -0.5f);
        const float fz = std::fmaf(tc.z, (float)s.nz, -0.5f);
        const float wx = fx - std::floor(fx), wy = fy - std::floor(fy), wz = fz - std::floor(fz);
        const int bx = glsl_floor_to_int(fx), by = glsl_floor_to_int(fy), bz = glsl_floor_to_int(fz);
        const int x0 = glsl_clampi(bx, 0, s.nx - 1), x1 = glsl_clampi(bx + 1, 0, s.nx - 1);
        const int y0 = glsl_clampi(by, 0, s.ny - 1), y1 = glsl_clampi(by + 1, 0, s.ny - 1);
        const int z0 = glsl_clampi(bz, 0, s.nz - 1), z1 = glsl_clampi(bz + 1, 0, s.nz - 1);
        const float v000 = s.texel(x0, y0, z0), v100 = s.texel(x1, y0, z0);
        const float v010 = s.texel(x0, y1, z0), v110 = s.texel(x1, y1, z0);
        const float v001 = s.texel(x0, y0, z1), v101 = s.texel(x1, y0, z1);
        const float v011 = s.texel(x0, y1, z1), v111 = s.texel(x1, y1, z1);
        const float c00 = std::fmaf(wx, v100 - v000, v000), c10 = std::fmaf(wx, v110 - v010, v010);
        const float c01 = std::fmaf(wx, v101 - v001, v001), c11 = std::fmaf(wx, v111 - v011, v011);
        const float c0 = std::fmaf(wy, c10 - c00, c00), c1 = std::fmaf(wy, c11 - c01, c01);
        r = std::fmaf(wz, c1 - c0, c0);
    }
    return utexel{r, 0.0f, 0.0f, 1.0f};
}

/* what every invocation sees besides the shader's own globals */
struct ShaderBase {
    uvec3 gl_GlobalInvocationID;
};
