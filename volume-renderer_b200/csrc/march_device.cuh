// march_device.cuh -- device-side ray march shared by every kernel variant.
//
// Bit-exactness contract: every arithmetic operation that VolumeRenderer.cs spells out is
// issued as exactly one correctly rounded binary32 instruction, in shader order, through
// the __f*_rn intrinsics (which nvcc never contracts into FMAs).  The only fused operations
// are the ones the trilinear EXTENSION is defined with (fma for the texel coordinate and
// the lerps) and the Markstein division sequence, which is checked on the
// device to return the correctly rounded quotient for the divisors in use
// (kernels_aux.cuh: verify_divisor_kernel).  The oracle (oracle/march_oracle.c) is the
// same sequence on the CPU; tests require max |diff| == 0.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "frame.h"

namespace vr {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// GLSL min/max: min(x,y) = y<x ? y : x ; max(x,y) = x<y ? y : x
__device__ __forceinline__ float glsl_min(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float glsl_max(float x, float y) { return (x < y) ? y : x; }

// a / d for a loop-invariant divisor d with inv = RN(1/d)
template <int MODE>
__device__ __forceinline__ float div_by(float a, float d, float inv)
{
    if (MODE == DIV_RECIP_EXACT) {
        return fmul(a, inv);
    } else if (MODE == DIV_MARKSTEIN) {
        const float q0 = fmul(a, inv);
        const float r = ffma(-d, q0, a);
        return ffma(r, inv, q0);
    } else {
        return fdiv(a, d);
    }
}

// exact integer -> float for v < 2^23 without the conversion pipe
__device__ __forceinline__ float u2f(uint32_t v)
{
    return fsub(__uint_as_float(0x4B000000u | v), 8388608.0f);
}

// Padded volume: (Nx+2) x (Ny+2) x (Nz+2) voxels, edge replicated, row pitch `pitch`
// elements, slice stride `slice` elements.  Padded index p holds voxel clamp(p-1, 0, N-1),
// which turns GL_CLAMP_TO_EDGE (RendererCore.cpp:411-413) into plain addressing.
template <typename T>
struct PaddedVolume {
    const T* __restrict__ base;
    uint32_t pitch;
    uint64_t slice;
    __device__ __forceinline__ const T* at(int jx, int jy, int jz) const
    {
        return base + ((uint64_t)(uint32_t)jz * slice + (uint64_t)((uint32_t)jy * pitch + (uint32_t)jx));
    }
};

// optional instrumentation: one bit per (unpadded) voxel
struct TouchMap {
    unsigned int* bits;       // nullptr when not counting
    int nx, ny, nz;
    __device__ __forceinline__ void mark(int jx, int jy, int jz) const
    {
        const int x = min(max(jx - 1, 0), nx - 1), y = min(max(jy - 1, 0), ny - 1), z = min(max(jz - 1, 0), nz - 1);
        const uint64_t idx = ((uint64_t)z * ny + y) * (uint64_t)nx + x;
        atomicOr(&bits[idx >> 5], 1u << (idx & 31));
    }
};

// VolumeRenderer.cs:121 -- nearest: i = clamp(floor(u*N), 0, N-1)  (padded: i+1, no clamp)
template <typename T, bool COUNT>
__device__ __forceinline__ float sample_nearest(const PaddedVolume<T>& vol, const FrameConsts& fc,
                                                float tx, float ty, float tz, const TouchMap& tm)
{
    const int jx = __float2int_rd(fmul(tx, fc.dimf[0])) + 1;
    const int jy = __float2int_rd(fmul(ty, fc.dimf[1])) + 1;
    const int jz = __float2int_rd(fmul(tz, fc.dimf[2])) + 1;
    if (COUNT) tm.mark(jx, jy, jz);
    return u2f((uint32_t)__ldg(vol.at(jx, jy, jz)));
}

// trilinear extension: f = fma(u,N,-0.5); i0 = floor(f); w = f - i0; lerp = fma(w, b-a, a)
template <typename T, bool COUNT>
__device__ __forceinline__ float sample_trilinear(const PaddedVolume<T>& vol, const FrameConsts& fc,
                                                  float tx, float ty, float tz, const TouchMap& tm)
{
    const float fx = ffma(tx, fc.dimf[0], -0.5f);
    const float fy = ffma(ty, fc.dimf[1], -0.5f);
    const float fz = ffma(tz, fc.dimf[2], -0.5f);
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const float wx = fsub(fx, flx), wy = fsub(fy, fly), wz = fsub(fz, flz);
    const int jx = __float2int_rd(fx) + 1, jy = __float2int_rd(fy) + 1, jz = __float2int_rd(fz) + 1;
    if (COUNT) {
        for (int c = 0; c < 8; ++c) tm.mark(jx + (c & 1), jy + ((c >> 1) & 1), jz + (c >> 2));
    }
    const T* p00 = vol.at(jx, jy, jz);
    const T* p10 = p00 + vol.pitch;
    const T* p01 = p00 + vol.slice;
    const T* p11 = p01 + vol.pitch;
    const float v000 = u2f(__ldg(p00)), v100 = u2f(__ldg(p00 + 1));
    const float v010 = u2f(__ldg(p10)), v110 = u2f(__ldg(p10 + 1));
    const float v001 = u2f(__ldg(p01)), v101 = u2f(__ldg(p01 + 1));
    const float v011 = u2f(__ldg(p11)), v111 = u2f(__ldg(p11 + 1));
    const float c00 = ffma(wx, fsub(v100, v000), v000);
    const float c10 = ffma(wx, fsub(v110, v010), v010);
    const float c01 = ffma(wx, fsub(v101, v001), v001);
    const float c11 = ffma(wx, fsub(v111, v011), v011);
    const float c0 = ffma(wy, fsub(c10, c00), c00);
    const float c1 = ffma(wy, fsub(c11, c01), c01);
    return ffma(wz, fsub(c1, c0), c0);
}

struct RaySetup {
    float org[3], dir[3];
    float t_min;
    bool hit;
};

// computeRay VolumeRenderer.cs:194-216 + intersectRayAABB :218-238 (per pixel, IEEE ops)
__device__ __forceinline__ RaySetup setup_ray(const FrameConsts& fc, int pix_x, int pix_y)
{
    RaySetup r;
    const float* cam = fc.cam;
    const float pixel_x = fadd((float)pix_x, 0.5f);
    const float pixel_y = fadd((float)pix_y, 0.5f);
    const float fw = (float)fc.W, fh = (float)fc.H;
    const float aspect = fdiv(fmul(fw, 1.0f), fh);
    const float x = fmul(aspect, fsub(fdiv(fmul(2.0f, pixel_x), fw), 1.0f));
    const float y = fsub(fdiv(fmul(2.0f, pixel_y), fh), 1.0f);
    const float z = -cam[20];
    float len = __fsqrt_rn(fadd(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)), fmul(0.0f, 0.0f)));
    const float dx = fdiv(x, len), dy = fdiv(y, len), dz = fdiv(z, len), dw = fdiv(0.0f, len);
    const float mx = fadd(fadd(fadd(fmul(cam[0], dx), fmul(cam[4], dy)), fmul(cam[8], dz)), fmul(cam[12], dw));
    const float my = fadd(fadd(fadd(fmul(cam[1], dx), fmul(cam[5], dy)), fmul(cam[9], dz)), fmul(cam[13], dw));
    const float mz = fadd(fadd(fadd(fmul(cam[2], dx), fmul(cam[6], dy)), fmul(cam[10], dz)), fmul(cam[14], dw));
    const float mw = fadd(fadd(fadd(fmul(cam[3], dx), fmul(cam[7], dy)), fmul(cam[11], dz)), fmul(cam[15], dw));
    len = __fsqrt_rn(fadd(fadd(fadd(fmul(mx, mx), fmul(my, my)), fmul(mz, mz)), fmul(mw, mw)));
    r.dir[0] = fdiv(mx, len); r.dir[1] = fdiv(my, len); r.dir[2] = fdiv(mz, len);
    r.org[0] = cam[16]; r.org[1] = cam[17]; r.org[2] = cam[18];

    float t_max = __int_as_float(0x7f800000), t_min = __int_as_float(0xff800000);
    float lo[3], hi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float inv = fdiv(1.0f, r.dir[i]);
        lo[i] = fmul(fsub(fc.pmin[i], r.org[i]), inv);
        hi[i] = fmul(fsub(fc.pmax[i], r.org[i]), inv);
    }
    t_min = glsl_max(t_min, glsl_min(lo[0], hi[0]));
    t_max = glsl_min(t_max, glsl_max(lo[0], hi[0]));
    t_min = glsl_max(t_min, glsl_min(lo[1], hi[1]));
    t_max = glsl_min(t_max, glsl_max(lo[1], hi[1]));
    if (t_max < t_min) {
        r.hit = false;
    } else {
        t_min = glsl_max(t_min, glsl_min(lo[2], hi[2]));
        t_max = glsl_min(t_max, glsl_max(lo[2], hi[2]));
        r.hit = (t_max > glsl_max(t_min, 0.0f));
    }
    r.t_min = t_min;
    return r;
}

// cartesianToTextureCoord, VolumeRenderer.cs:175-192
template <int TCDIV, bool GENERIC>
__device__ __forceinline__ void tex_coord(const FrameConsts& fc, float px, float py, float pz,
                                          float& tx, float& ty, float& tz)
{
    const float qx = div_by<TCDIV>(fadd(px, fc.half_len[0]), fc.denom[0], fc.inv_denom[0]);
    const float qy = div_by<TCDIV>(fadd(py, fc.half_len[1]), fc.denom[1], fc.inv_denom[1]);
    float qz = div_by<TCDIV>(fadd(pz, fc.half_len[2]), fc.denom[2], fc.inv_denom[2]);
    qz = fsub(1.0f, qz);
    if (GENERIC && fc.view_top) { tx = qx; ty = fsub(1.0f, qz); tz = qy; }
    else if (GENERIC && fc.view_bottom) { tx = qx; ty = qz; tz = fsub(1.0f, qy); }
    else { tx = qx; ty = qy; tz = qz; }
}

struct MarchCounters { unsigned long long samples; };

// rayMarchVolume VolumeRenderer.cs:104-139 / MIP :141-173 for one ray whose setup hit the box.
//   T       voxel type            FILTER  VR_FILTER_* (ignored when GENERIC: runtime fc.filter)
//   TCDIV   tex-coord division    GENERIC runtime handling of MIP / TF / view swizzle /
//                                         opacity correction / unordered window (IEEE window div)
template <typename T, int FILTER, int TCDIV, bool GENERIC, bool COUNT>
__device__ __forceinline__ void march_ray(const FrameConsts& fc, const PaddedVolume<T>& vol,
                                          const float* __restrict__ tf_lut, const RaySetup& r,
                                          const TouchMap& tm, float& outC, float& outA,
                                          unsigned long long& nsamples)
{
    const float EPSILON = 0.000001f;
    float pos[3], dstep[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float start = fadd(r.org[i], fmul(r.dir[i], r.t_min));   // :107
        pos[i] = fadd(start, fmul(r.dir[i], EPSILON));                  // :114
        dstep[i] = fmul(r.dir[i], fc.step);                             // :136
    }
    float C = 0.0f, A = 0.0f;
    for (int i = 0; i < 10000; ++i) {
        float tx, ty, tz;
        tex_coord<TCDIV, GENERIC>(fc, pos[0], pos[1], pos[2], tx, ty, tz);
        if (tx > 1.0f || ty > 1.0f || tz > 1.0f || tx < 0.0f || ty < 0.0f || tz < 0.0f || A >= 0.95f)
            break;                                                      // :118

        float s;
        if (GENERIC ? (fc.filter == VR_FILTER_NEAREST) : (FILTER == VR_FILTER_NEAREST))
            s = sample_nearest<T, COUNT>(vol, fc, tx, ty, tz, tm);
        else
            s = sample_trilinear<T, COUNT>(vol, fc, tx, ty, tz, tm);
        if (COUNT) ++nsamples;

        float v;
        if (GENERIC) {
            v = glsl_min(glsl_max(s, fc.fmin), fc.fmax);                // :122
            if (v <= fc.fmax && v >= fc.fmin)                           // :123
                v = fdiv(fsub(v, fc.fmin), fc.frange);                  // :124
        } else {
            v = fminf(fmaxf(s, fc.fmin), fc.fmax);
            v = div_by<DIV_MARKSTEIN>(fsub(v, fc.fmin), fc.frange, fc.inv_frange);
        }

        float src_rgb = v, src_a = v;
        if (GENERIC && fc.use_tf) {
            int iso = __float2int_rd(fadd(fmul(v, 255.0f), 0.5f));
            iso = min(max(iso, 0), 255);
            src_a = __ldg(tf_lut + iso);
        }
        if (GENERIC && fc.is_mip) {
            src_rgb = fmul(src_rgb, fc.alpha_scale);                    // :163
            src_a = fmul(src_a, fc.alpha_scale);
            if (A < src_a) { C = src_rgb; A = src_a; }                  // :164-167
        } else {
            src_a = fmul(src_a, fc.alpha_scale);                        // :130
            if (GENERIC && fc.opacity_correction)
                src_a = (float)(1.0 - pow(1.0 - (double)src_a, (double)fc.step_scale));
            src_rgb = fmul(src_rgb, src_a);                             // :131
            const float t = fsub(1.0f, A);                              // :132
            C = fadd(C, fmul(src_rgb, t));
            A = fadd(A, fmul(src_a, t));
            if (A > 0.99f) break;                                       // :134
        }
        pos[0] = fadd(pos[0], dstep[0]);                                // :136
        pos[1] = fadd(pos[1], dstep[1]);
        pos[2] = fadd(pos[2], dstep[2]);
    }
    outC = C; outA = A;
}

// local (owned) row -> global image row under the screen-row-tile partition
__device__ __forceinline__ int owned_row_to_global(const FrameConsts& fc, int local_row)
{
    const int tile_local = local_row / fc.tile_rows;
    const int r_in = local_row - tile_local * fc.tile_rows;
    return (tile_local * fc.world + fc.rank) * fc.tile_rows + r_in;
}

}  // namespace vr
