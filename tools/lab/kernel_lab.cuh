// kernel_lab.cuh (round 1: csrc/kernel_fast.cuh) -- DEVELOPMENT KERNELS, not part of the product library.
// issue-slot-optimised form of the reference march for the common case
// (DVR, default view, no transfer function, ordered window, alpha_scale >= 0).
//
// Same correctly-rounded operation sequence as march_device.cuh / the oracle -- results are
// bit-identical -- but arranged for the B200 SM, where this loop is instruction-issue bound
// (profiles/r01_*): two rays (horizontally adjacent pixels) per thread with every IEEE
// add/mul/fma issued as ONE packed Blackwell f32x2 instruction (FADD2/FMUL2/FFMA2) for both
// rays; texel fetch from the edge-replicated volume without clamps; u16->f32 through the
// exponent-bias trick with the x-differences taken on the biased values (exact); range tests
// on the float bit patterns with 3-input integer max; divisions by loop-invariant divisors
// through the device-verified Markstein sequence.
#pragma once

#include "f32x2.cuh"
#include "march_device.cuh"

namespace vr {

// ptxas (12.9) contracts mul.rn.f32x2 -> add/sub.rn.f32x2 into FFMA2 even under -fmad=false
// (it honours the explicit .rn only on scalar ops).  Wherever the shader has an UNFUSED
// product feeding a sum, the sum is therefore issued as two scalar add.rn.f32.
__device__ __forceinline__ f2 fadd_after_mul(f2 a, f2 prod) { return mk2(__fadd_rn(lo(a), lo(prod)), __fadd_rn(hi(a), hi(prod))); }
__device__ __forceinline__ f2 fsub_after_mul(f2 a, f2 prod) { return mk2(__fsub_rn(lo(a), lo(prod)), __fsub_rn(hi(a), hi(prod))); }

// scalar "pair" of one ray so the same march body serves 1 and 2 rays per thread
struct f1 { float v; };
__device__ __forceinline__ f1 fadd(f1 a, f1 b) { return {__fadd_rn(a.v, b.v)}; }
__device__ __forceinline__ f1 fsub(f1 a, f1 b) { return {__fsub_rn(a.v, b.v)}; }
__device__ __forceinline__ f1 fmul(f1 a, f1 b) { return {__fmul_rn(a.v, b.v)}; }
__device__ __forceinline__ f1 ffma(f1 a, f1 b, f1 c) { return {__fmaf_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ f1 fadd_after_mul(f1 a, f1 prod) { return {__fadd_rn(a.v, prod.v)}; }
__device__ __forceinline__ f1 fsub_after_mul(f1 a, f1 prod) { return {__fsub_rn(a.v, prod.v)}; }

template <typename L> struct Lanes;
template <> struct Lanes<f1> {
    static constexpr int N = 1;
    static __device__ __forceinline__ f1 splat(float a) { return {a}; }
    static __device__ __forceinline__ float get(f1 p, int) { return p.v; }
    template <typename F> static __device__ __forceinline__ f1 map(F f) { return {f(0)}; }
};
template <> struct Lanes<f2> {
    static constexpr int N = 2;
    static __device__ __forceinline__ f2 splat(float a) { return splat2(a); }
    static __device__ __forceinline__ float get(f2 p, int i) { return i == 0 ? lo(p) : hi(p); }
    template <typename F> static __device__ __forceinline__ f2 map(F f) { return mk2(f(0), f(1)); }
};

enum WinMode : int {
    WIN_CLAMP = 0,     // clamp to [min,max], subtract min, divide by range (any ordered window)
    WIN_COVERS0 = 1    // min == 0 and every voxel value <= max: clamp and subtraction are no-ops
};

struct FastArgs {
    const void* vol;
    uint32_t pitch;            // elements per row of the padded volume
    uint32_t slice_lo;         // pitch * (Ny+2), elements (< 2^32 checked on the host)
    float* out;
    int local_rows;
};

// a / d, d loop invariant, packed or scalar
template <int MODE, typename L>
__device__ __forceinline__ L div_by_l(L a, float d, float inv)
{
    const L vinv = Lanes<L>::splat(inv);
    if (MODE == DIV_RECIP_EXACT) return fmul(a, vinv);
    const L q0 = fmul(a, vinv);
    const L r = ffma(Lanes<L>::splat(-d), q0, a);
    return ffma(r, vinv, q0);
}

enum FloorMode : int {
    FLOOR_XU2 = 0,     // FRND.FLOOR + F2I.FLOOR            (2 conversion-pipe ops per axis)
    FLOOR_XU1 = 1,     // F2I.FLOOR + I2FP                   (1 conversion-pipe op per axis)
    FLOOR_MAGIC = 2    // 1.5*2^23 rounding trick            (no conversion-pipe op)
};

// f -> (w = f - floor(f), floor(f) + 1 as int), all lanes.  Every variant returns exactly
// floorf(f) and the single-rounded difference f - floorf(f) for |f| < 2^22.
template <int FM, typename L>
__device__ __forceinline__ void floor_frac_idx(L f, L& w, int idx_plus1[2])
{
    typedef Lanes<L> LN;
    if (FM == FLOOR_XU2) {
        const L fl = LN::map([&](int l) { return floorf(LN::get(f, l)); });
        w = fsub(f, fl);
#pragma unroll
        for (int l = 0; l < LN::N; ++l) idx_plus1[l] = __float2int_rd(LN::get(f, l)) + 1;
    } else if (FM == FLOOR_XU1) {
#pragma unroll
        for (int l = 0; l < LN::N; ++l) idx_plus1[l] = __float2int_rd(LN::get(f, l));
        const L fl = LN::map([&](int l) { return (float)idx_plus1[l]; });
        w = fsub(f, fl);
#pragma unroll
        for (int l = 0; l < LN::N; ++l) idx_plus1[l] += 1;
    } else {
        const L M = LN::splat(12582912.0f);                // 1.5 * 2^23: t = M + RN(f) exactly
        const L t = fadd(f, M);
        const L r = fsub(t, M);                             // RN(f)
        bool up[2];
#pragma unroll
        for (int l = 0; l < LN::N; ++l) up[l] = LN::get(r, l) > LN::get(f, l);   // rounded up: floor = RN(f) - 1
        const L adj = LN::map([&](int l) { return up[l] ? 1.0f : 0.0f; });
        w = fsub(f, fsub(r, adj));
#pragma unroll
        for (int l = 0; l < LN::N; ++l)
            idx_plus1[l] = __float_as_int(LN::get(t, l)) - 0x4B400000 + (up[l] ? 0 : 1);
    }
}

// VolumeRenderer.cs:121 for all lanes: nearest i = floor(u*N) (+1 padded, no clamp needed), or
// the trilinear extension f = fma(u,N,-0.5), lerp(a,b,w) = fma(w, b-a, a), x then y then z.
// pair-packed layout: element e of the "pairs" volume holds voxels (x, x+1) of the padded
// volume in one word (uint32 for 16-bit data, uint16 for 8-bit data): one load per texel row
template <typename T> struct PairWord;
template <> struct PairWord<uint16_t> { typedef uint32_t type; };
template <> struct PairWord<uint8_t>  { typedef uint16_t type; };
template <typename T> __device__ __forceinline__ void unpack_biased(typename PairWord<T>::type w, float& lo_b, float& hi_b);
template <> __device__ __forceinline__ void unpack_biased<uint16_t>(uint32_t w, float& lo_b, float& hi_b)
{
    lo_b = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610));
    hi_b = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632));
}
template <> __device__ __forceinline__ void unpack_biased<uint8_t>(uint16_t w, float& lo_b, float& hi_b)
{
    lo_b = __uint_as_float(__byte_perm((uint32_t)w, 0x4B000000u, 0x7660));
    hi_b = __uint_as_float(__byte_perm((uint32_t)w, 0x4B000000u, 0x7661));
}

template <typename T, int FILTER, int FM, bool PAIRS, typename L>
__device__ __forceinline__ L sample_lanes(const T* __restrict__ vol, uint32_t pitch, uint32_t slice,
                                          const float dimf[3], L tx, L ty, L tz)
{
    typedef Lanes<L> LN;
    if (FILTER == VR_FILTER_NEAREST) {
        const L ux = fmul(tx, LN::splat(dimf[0])), uy = fmul(ty, LN::splat(dimf[1])), uz = fmul(tz, LN::splat(dimf[2]));
        return LN::map([&](int l) {
            const int jx = __float2int_rd(LN::get(ux, l)) + 1, jy = __float2int_rd(LN::get(uy, l)) + 1,
                      jz = __float2int_rd(LN::get(uz, l)) + 1;
            const uint32_t e = (uint32_t)jz * slice + ((uint32_t)jy * pitch + (uint32_t)jx);
            return u2f((uint32_t)__ldg(vol + e));
        });
    }
    const L mhalf = LN::splat(-0.5f);
    L wx, wy, wz;
    int jx[2], jy[2], jz[2];
    floor_frac_idx<FM>(ffma(tx, LN::splat(dimf[0]), mhalf), wx, jx);
    floor_frac_idx<FM>(ffma(ty, LN::splat(dimf[1]), mhalf), wy, jy);
    floor_frac_idx<FM>(ffma(tz, LN::splat(dimf[2]), mhalf), wz, jz);
    // biased floats 2^23 + v: differences of biased values are exact, so only the four
    // x-low corners need un-biasing
    float b[2][8];
#pragma unroll
    for (int l = 0; l < LN::N; ++l) {
        // 32-bit element index: the host only selects this kernel when the padded volume has
        // fewer than 2^32 voxels
        const uint32_t e00 = (uint32_t)jz[l] * slice + ((uint32_t)jy[l] * pitch + (uint32_t)jx[l]);
        const uint32_t e10 = e00 + pitch, e01 = e00 + slice, e11 = e01 + pitch;
        if (PAIRS) {
            const typename PairWord<T>::type* pv = reinterpret_cast<const typename PairWord<T>::type*>(vol);
            unpack_biased<T>(__ldg(pv + e00), b[l][0], b[l][1]);
            unpack_biased<T>(__ldg(pv + e10), b[l][2], b[l][3]);
            unpack_biased<T>(__ldg(pv + e01), b[l][4], b[l][5]);
            unpack_biased<T>(__ldg(pv + e11), b[l][6], b[l][7]);
            continue;
        }
#if defined(VR_ABLATE) && VR_ABLATE == 1      // lab only: no memory traffic at all
        for (int k = 0; k < 8; ++k) b[l][k] = __uint_as_float(0x4B000000u | ((e00 + 37u * k) & 0xfffu));
        continue;
#elif defined(VR_ABLATE) && VR_ABLATE == 2    // lab only: every lane reads the same 8 texels (broadcast)
        { const uint32_t u = (e00 & 0u) + 1234567u, u10 = u + pitch, u01 = u + slice, u11 = u01 + pitch;
          b[l][0] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u)); b[l][1] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u + 1));
          b[l][2] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u10)); b[l][3] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u10 + 1));
          b[l][4] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u01)); b[l][5] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u01 + 1));
          b[l][6] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u11)); b[l][7] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + u11 + 1));
          continue; }
#endif
        b[l][0] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e00));
        b[l][1] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e00 + 1));
        b[l][2] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e10));
        b[l][3] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e10 + 1));
        b[l][4] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e01));
        b[l][5] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e01 + 1));
        b[l][6] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e11));
        b[l][7] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e11 + 1));
    }
    L v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = LN::map([&](int l) { return b[l][k]; });
    const L B = LN::splat(8388608.0f);
    const L c00 = ffma(wx, fsub(v[1], v[0]), fsub(v[0], B));
    const L c10 = ffma(wx, fsub(v[3], v[2]), fsub(v[2], B));
    const L c01 = ffma(wx, fsub(v[5], v[4]), fsub(v[4], B));
    const L c11 = ffma(wx, fsub(v[7], v[6]), fsub(v[6], B));
    const L c0 = ffma(wy, fsub(c10, c00), c00);
    const L c1 = ffma(wy, fsub(c11, c01), c01);
    return ffma(wz, fsub(c1, c0), c0);
}

// One march loop for NL = Lanes<L>::N rays advancing in lock step.  Returns with `done` bits
// set for rays that have finished (left the box or reached the opacity threshold); a ray
// whose partner finished first is completed by a scalar call of the same function.
template <typename T, int FILTER, int TCDIV, int WIN, int FM, typename L, bool PAIRS = false>
__device__ __forceinline__ unsigned march_lanes(const FrameConsts& fc, const T* __restrict__ vol, uint32_t pitch,
                                                uint32_t slice, L pos[3], const L dstep[3], L& C, L& A, int& iter)
{
    typedef Lanes<L> LN;
    const L half0 = LN::splat(fc.half_len[0]), half1 = LN::splat(fc.half_len[1]), half2 = LN::splat(fc.half_len[2]);
    const L one = LN::splat(1.0f);
    const L alpha = LN::splat(fc.alpha_scale);
    const L vfmin = LN::splat(fc.fmin);
    unsigned done = 0;
    for (; iter < 10000; ++iter) {
        // cartesianToTextureCoord, VolumeRenderer.cs:175-192
        const L tx = div_by_l<TCDIV>(fadd(pos[0], half0), fc.denom[0], fc.inv_denom[0]);
        const L ty = div_by_l<TCDIV>(fadd(pos[1], half1), fc.denom[1], fc.inv_denom[1]);
        const L tz = fsub_after_mul(one, div_by_l<TCDIV>(fadd(pos[2], half2), fc.denom[2], fc.inv_denom[2]));
        // :118 -- 0 <= t <= 1 on all three and A < 0.95, on the bit patterns (no -0, no NaN:
        // the camera block is validated finite and alpha_scale >= 0 on this path)
#pragma unroll
        for (int l = 0; l < LN::N; ++l) {
            const unsigned m = max(max(__float_as_uint(LN::get(tx, l)), __float_as_uint(LN::get(ty, l))),
                                   __float_as_uint(LN::get(tz, l)));
            if (m > 0x3F800000u || __float_as_uint(LN::get(A, l)) >= 0x3F733333u) done |= 1u << l;
        }
        if (done) break;

        const L s = sample_lanes<T, FILTER, FM, PAIRS>(vol, pitch, slice, fc.dimf, tx, ty, tz);
        // :122-124
        L v;
        if (WIN == WIN_COVERS0) {
            v = div_by_l<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        } else {
            const L cl = LN::map([&](int l) { return fminf(fmaxf(LN::get(s, l), fc.fmin), fc.fmax); });
            v = div_by_l<DIV_MARKSTEIN>(fsub(cl, vfmin), fc.frange, fc.inv_frange);
        }
        // :130-132 (the bottom `dest.a > 0.99` break of :134 is subsumed by the :118 test of
        // the next iteration: nothing but `pos` changes in between)
        const L a = fmul(v, alpha);
        const L c = fmul(v, a);
        const L t = fsub(one, A);
        C = fadd_after_mul(C, fmul(c, t));
        A = fadd_after_mul(A, fmul(a, t));
        pos[0] = fadd(pos[0], dstep[0]);
        pos[1] = fadd(pos[1], dstep[1]);
        pos[2] = fadd(pos[2], dstep[2]);
    }
    return done;
}

// ------------------------------------------------------------------------------------------
// Software-pipelined march.  The texel addresses of sample i+1 depend only on `pos`, never on
// loaded data; the only data dependence between iterations is the opacity test.  So the loads
// of sample i+1 are issued BEFORE sample i is interpolated and composited: two samples (four
// with 2 rays/thread) are in flight per thread, which is what hides the L1-miss latency this
// loop is otherwise bound by.  Scheduling only -- the operation sequence per sample, and hence
// every bit of the result, is unchanged.  A prefetch that turns out to be unnecessary (the ray
// terminated on opacity) is a harmless read inside the padded volume.
template <typename T, bool PAIRS> struct RawTexels { uint32_t w[PAIRS ? 4 : 8]; };

template <typename T, int FM, bool PAIRS, typename L>
struct Fetch {
    L wx, wy, wz;
    RawTexels<T, PAIRS> raw[2];
};

// coordinates -> weights + issued loads (trilinear only)
template <typename T, int FM, bool PAIRS, typename L>
__device__ __forceinline__ void issue_fetch(const T* __restrict__ vol, uint32_t pitch, uint32_t slice, const float dimf[3],
                                            L tx, L ty, L tz, Fetch<T, FM, PAIRS, L>& f)
{
    typedef Lanes<L> LN;
    const L mhalf = LN::splat(-0.5f);
    int jx[2], jy[2], jz[2];
    floor_frac_idx<FM>(ffma(tx, LN::splat(dimf[0]), mhalf), f.wx, jx);
    floor_frac_idx<FM>(ffma(ty, LN::splat(dimf[1]), mhalf), f.wy, jy);
    floor_frac_idx<FM>(ffma(tz, LN::splat(dimf[2]), mhalf), f.wz, jz);
#pragma unroll
    for (int l = 0; l < LN::N; ++l) {
        const uint32_t e00 = (uint32_t)jz[l] * slice + ((uint32_t)jy[l] * pitch + (uint32_t)jx[l]);
        const uint32_t e10 = e00 + pitch, e01 = e00 + slice, e11 = e01 + pitch;
        if (PAIRS) {
            const typename PairWord<T>::type* pv = reinterpret_cast<const typename PairWord<T>::type*>(vol);
            f.raw[l].w[0] = __ldg(pv + e00); f.raw[l].w[1] = __ldg(pv + e10);
            f.raw[l].w[2] = __ldg(pv + e01); f.raw[l].w[3] = __ldg(pv + e11);
        } else {
            f.raw[l].w[0] = __ldg(vol + e00); f.raw[l].w[1] = __ldg(vol + e00 + 1);
            f.raw[l].w[2] = __ldg(vol + e10); f.raw[l].w[3] = __ldg(vol + e10 + 1);
            f.raw[l].w[4] = __ldg(vol + e01); f.raw[l].w[5] = __ldg(vol + e01 + 1);
            f.raw[l].w[6] = __ldg(vol + e11); f.raw[l].w[7] = __ldg(vol + e11 + 1);
        }
    }
}

// loaded texels + weights -> interpolated sample (same arithmetic as sample_lanes)
template <typename T, int FM, bool PAIRS, typename L>
__device__ __forceinline__ L finish_fetch(const Fetch<T, FM, PAIRS, L>& f)
{
    typedef Lanes<L> LN;
    float b[2][8];
#pragma unroll
    for (int l = 0; l < LN::N; ++l) {
        if (PAIRS) {
#pragma unroll
            for (int k = 0; k < 4; ++k) unpack_biased<T>((typename PairWord<T>::type)f.raw[l].w[k], b[l][2 * k], b[l][2 * k + 1]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) b[l][k] = __uint_as_float(0x4B000000u | f.raw[l].w[k]);
        }
    }
    L v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = LN::map([&](int l) { return b[l][k]; });
    const L B = LN::splat(8388608.0f);
    const L c00 = ffma(f.wx, fsub(v[1], v[0]), fsub(v[0], B));
    const L c10 = ffma(f.wx, fsub(v[3], v[2]), fsub(v[2], B));
    const L c01 = ffma(f.wx, fsub(v[5], v[4]), fsub(v[4], B));
    const L c11 = ffma(f.wx, fsub(v[7], v[6]), fsub(v[6], B));
    const L c0 = ffma(f.wy, fsub(c10, c00), c00);
    const L c1 = ffma(f.wy, fsub(c11, c01), c01);
    return ffma(f.wz, fsub(c1, c0), c0);
}

// Same contract as march_lanes (trilinear): `pos` is the position of the next sample to take on
// entry and on exit; returns the `done` bits.
template <typename T, int TCDIV, int WIN, int FM, typename L, bool PAIRS>
__device__ __forceinline__ unsigned march_lanes_pipe(const FrameConsts& fc, const T* __restrict__ vol, uint32_t pitch,
                                                     uint32_t slice, L pos[3], const L dstep[3], L& C, L& A, int& iter)
{
    typedef Lanes<L> LN;
    const L half0 = LN::splat(fc.half_len[0]), half1 = LN::splat(fc.half_len[1]), half2 = LN::splat(fc.half_len[2]);
    const L one = LN::splat(1.0f);
    const L alpha = LN::splat(fc.alpha_scale);
    const L vfmin = LN::splat(fc.fmin);
    auto coords = [&](const L p[3], L& tx, L& ty, L& tz) -> unsigned {          // :175-192 + the box part of :118
        tx = div_by_l<TCDIV>(fadd(p[0], half0), fc.denom[0], fc.inv_denom[0]);
        ty = div_by_l<TCDIV>(fadd(p[1], half1), fc.denom[1], fc.inv_denom[1]);
        tz = fsub_after_mul(one, div_by_l<TCDIV>(fadd(p[2], half2), fc.denom[2], fc.inv_denom[2]));
        unsigned out = 0;
#pragma unroll
        for (int l = 0; l < LN::N; ++l) {
            const unsigned m = max(max(__float_as_uint(LN::get(tx, l)), __float_as_uint(LN::get(ty, l))), __float_as_uint(LN::get(tz, l)));
            if (m > 0x3F800000u) out |= 1u << l;
        }
        return out;
    };
    auto opaque = [&]() -> unsigned {                                            // the opacity part of :118
        unsigned o = 0;
#pragma unroll
        for (int l = 0; l < LN::N; ++l) if (__float_as_uint(LN::get(A, l)) >= 0x3F733333u) o |= 1u << l;
        return o;
    };

    L tx, ty, tz;
    unsigned done = coords(pos, tx, ty, tz) | opaque();
    if (done || iter >= 10000) return done;
    Fetch<T, FM, PAIRS, L> cur;
    issue_fetch<T, FM, PAIRS, L>(vol, pitch, slice, fc.dimf, tx, ty, tz, cur);
    for (;;) {
        // stage A for the next sample: position, box test, addresses, loads
        L npos[3] = {fadd(pos[0], dstep[0]), fadd(pos[1], dstep[1]), fadd(pos[2], dstep[2])};     // :136
        const unsigned out_next = coords(npos, tx, ty, tz);
        Fetch<T, FM, PAIRS, L> nxt;
        if (!out_next) issue_fetch<T, FM, PAIRS, L>(vol, pitch, slice, fc.dimf, tx, ty, tz, nxt);
        // stage B for the pending sample: interpolate, window, composite
        const L s = finish_fetch<T, FM, PAIRS, L>(cur);
        L v;
        if (WIN == WIN_COVERS0) {
            v = div_by_l<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        } else {
            const L cl = LN::map([&](int l) { return fminf(fmaxf(LN::get(s, l), fc.fmin), fc.fmax); });
            v = div_by_l<DIV_MARKSTEIN>(fsub(cl, vfmin), fc.frange, fc.inv_frange);
        }
        const L a = fmul(v, alpha);                                              // :130-132
        const L c = fmul(v, a);
        const L t = fsub(one, A);
        C = fadd_after_mul(C, fmul(c, t));
        A = fadd_after_mul(A, fmul(a, t));
        ++iter;
        pos[0] = npos[0]; pos[1] = npos[1]; pos[2] = npos[2];
        done = out_next | opaque();
        if (done || iter >= 10000) break;
        cur = nxt;
    }
    return done;
}

// ------------------------------------------------------------------------------------------
// One ray per thread with f32x2 packing INSIDE the ray: the x/y components of position, texture
// coordinate, texel coordinate and weights travel as one packed pair, the four x-lerps run as
// two packed lerps over the (z, z+1) row pairs, the two y-lerps as one, and c*t / a*t as one
// packed multiply.  ~14 fewer issue slots per sample than the scalar loop at the same register
// count and occupancy.  Same operation sequence per value as march_lanes -> identical bits.
// UNIT: every tex-coord divisor is exactly 1 (cubic volume, unit spacing): q/1 = q, no multiply.
// NOCAP: the host proved the 10000-iteration cap of VolumeRenderer.cs:115 cannot be reached.
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP>
__device__ __forceinline__ void march_ray_packed(const FrameConsts& fc, const T* __restrict__ vol, uint32_t pitch, uint32_t slice,
                                                 const float pos0[3], const float dstep[3], float& outC, float& outA)
{
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 hxy = mk2(fc.half_len[0], fc.half_len[1]);
    const float hz = fc.half_len[2];
    const f2 ixy = mk2(fc.inv_denom[0], fc.inv_denom[1]);
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const f2 mhalf = splat2(-0.5f), B2 = splat2(8388608.0f);
    const T* __restrict__ volp = vol + ((size_t)slice + pitch + 1);
    float C = 0.0f, A = 0.0f;
    for (int iter = 0; NOCAP || iter < 10000; ++iter) {
        // cartesianToTextureCoord :175-192
        const f2 qxy = fadd(pxy, hxy);
        const float qz = __fadd_rn(pz, hz);
        f2 txy;
        float tzq;
        if (UNIT) { txy = qxy; tzq = qz; }
        else if (TCDIV == DIV_RECIP_EXACT) { txy = fmul(qxy, ixy); tzq = __fmul_rn(qz, fc.inv_denom[2]); }
        else {
            const f2 q0 = fmul(qxy, ixy);
            const f2 r = ffma(mk2(-fc.denom[0], -fc.denom[1]), q0, qxy);
            txy = ffma(r, ixy, q0);
            tzq = div_by<DIV_MARKSTEIN>(qz, fc.denom[2], fc.inv_denom[2]);
        }
        const float tz = __fsub_rn(1.0f, tzq);
        const unsigned m = max(max(__float_as_uint(lo(txy)), __float_as_uint(hi(txy))), __float_as_uint(tz));
        if (m > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) break;           // :118
        // texel coordinates, indices, weights
        const f2 fxy = ffma(txy, nxy, mhalf);
        const float fz = __fmaf_rn(tz, nz, -0.5f);
        const int ix = __float2int_rd(lo(fxy)), iy = __float2int_rd(hi(fxy)), iz = __float2int_rd(fz);
        const f2 wxy = fsub(fxy, mk2((float)ix, (float)iy));
        const float wz = __fsub_rn(fz, (float)iz);
        // texel (ix+1, iy+1, iz+1) of the padded volume; the +1s live in `volp` (signed 32-bit index:
        // the host selects this kernel only below 2^31 padded voxels)
        const int e00 = iz * (int)slice + (iy * (int)pitch + ix);
        const int e10 = e00 + (int)pitch, e01 = e00 + (int)slice, e11 = e01 + (int)pitch;
        // biased texels 2^23 + v, paired over (z, z+1): A = rows (y,z),(y,z+1); B = rows (y+1,z),(y+1,z+1)
        const f2 loA = mk2(__uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e00)),     __uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e01)));
        const f2 hiA = mk2(__uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e00 + 1)), __uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e01 + 1)));
        const f2 loB = mk2(__uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e10)),     __uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e11)));
        const f2 hiB = mk2(__uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e10 + 1)), __uint_as_float(0x4B000000u | (uint32_t)__ldg(volp + e11 + 1)));
        const f2 wxx = splat2(lo(wxy)), wyy = splat2(hi(wxy));
        const f2 cA = ffma(wxx, fsub(hiA, loA), fsub(loA, B2));        // (c00, c01)
        const f2 cB = ffma(wxx, fsub(hiB, loB), fsub(loB, B2));        // (c10, c11)
        const f2 cy = ffma(wyy, fsub(cB, cA), cA);                     // (c0, c1)
        const float s = __fmaf_rn(wz, __fsub_rn(hi(cy), lo(cy)), lo(cy));
        // :122-124
        float v;
        if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fc.fmin), fc.fmax), fc.fmin), fc.frange, fc.inv_frange);
        // :130-132
        const float a = __fmul_rn(v, fc.alpha_scale);
        const float c = __fmul_rn(v, a);
        const float t = __fsub_rn(1.0f, A);
        const f2 ca_t = fmul(mk2(c, a), splat2(t));
        C = __fadd_rn(C, lo(ca_t));
        A = __fadd_rn(A, hi(ca_t));
        pxy = fadd(pxy, dxy);                                          // :136
        pz = __fadd_rn(pz, dz);
    }
    outC = C; outA = A;
}

// ------------------------------------------------------------------------------------------
// Texture-gather variant of the packed march.  The volume additionally lives in a layered 2-D
// CUDA array (layer = z slice, point sampling, clamp-to-edge, element read mode).  One `tld4`
// returns the four texels of the bilinear footprint of one slice -- exact integer values, no
// filtering arithmetic in the texture unit -- so a sample costs 2 texture instructions instead
// of 8 loads + 9 address instructions, on a pipe (TEX) that is otherwise idle.  The gather
// coordinate is the CORNER shared by the four texels (ix+1, iy+1): half a texel away from
// every footprint boundary, hence immune to the unit's fixed-point coordinate rounding.
// Interpolation, windowing and compositing stay in fp32 ALU exactly as in march_ray_packed.
__device__ __forceinline__ void tld4_layer(cudaTextureObject_t tex, int layer, float x, float y,
                                           uint32_t& t_x0y1, uint32_t& t_x1y1, uint32_t& t_x1y0, uint32_t& t_x0y0)
{
    asm volatile("tld4.r.a2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                 : "=r"(t_x0y1), "=r"(t_x1y1), "=r"(t_x1y0), "=r"(t_x0y0)
                 : "l"(tex), "r"(layer), "f"(x), "f"(y));
}

__device__ __forceinline__ void tld4_layer_f32(cudaTextureObject_t tex, int layer, float x, float y,
                                               float& t_x0y1, float& t_x1y1, float& t_x1y0, float& t_x0y0)
{
    asm volatile("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                 : "=f"(t_x0y1), "=f"(t_x1y1), "=f"(t_x1y0), "=f"(t_x0y0)
                 : "l"(tex), "r"(layer), "f"(x), "f"(y));
}

// FLOATTEX (lab experiment): the array holds float(v) texels, so the gather returns exact floats and
// the 8 exponent-bias conversions + 4 un-bias subtractions disappear (at 2x array bytes)
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, bool FLOATTEX = false>
__device__ __forceinline__ void march_ray_texgather(const FrameConsts& fc, cudaTextureObject_t tex,
                                                    const float pos0[3], const float dstep[3], float& outC, float& outA)
{
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 hxy = mk2(fc.half_len[0], fc.half_len[1]);
    const float hz = fc.half_len[2];
    const f2 ixy = mk2(fc.inv_denom[0], fc.inv_denom[1]);
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const int zmax = fc.dim[2] - 1;
    const f2 mhalf = splat2(-0.5f), B2 = splat2(8388608.0f), one2 = splat2(1.0f);
    float C = 0.0f, A = 0.0f;
    for (int iter = 0; NOCAP || iter < 10000; ++iter) {
        const f2 qxy = fadd(pxy, hxy);
        const float qz = __fadd_rn(pz, hz);
        f2 txy;
        float tzq;
        if (UNIT) { txy = qxy; tzq = qz; }
        else if (TCDIV == DIV_RECIP_EXACT) { txy = fmul(qxy, ixy); tzq = __fmul_rn(qz, fc.inv_denom[2]); }
        else {
            const f2 q0 = fmul(qxy, ixy);
            const f2 r = ffma(mk2(-fc.denom[0], -fc.denom[1]), q0, qxy);
            txy = ffma(r, ixy, q0);
            tzq = div_by<DIV_MARKSTEIN>(qz, fc.denom[2], fc.inv_denom[2]);
        }
        const float tz = __fsub_rn(1.0f, tzq);
        const unsigned m = max(max(__float_as_uint(lo(txy)), __float_as_uint(hi(txy))), __float_as_uint(tz));
        if (m > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) break;           // :118
        const f2 fxy = ffma(txy, nxy, mhalf);
        const float fz = __fmaf_rn(tz, nz, -0.5f);
        const int ix = __float2int_rd(lo(fxy)), iy = __float2int_rd(hi(fxy)), iz = __float2int_rd(fz);
        const f2 flxy = mk2((float)ix, (float)iy);
        const f2 wxy = fsub(fxy, flxy);
        const float wz = __fsub_rn(fz, (float)iz);
        const f2 cxy = fadd(flxy, one2);                               // corner shared by the 2x2 footprint
        const f2 wxx = splat2(lo(wxy)), wyy = splat2(hi(wxy));
        f2 cA, cB;
        if (FLOATTEX) {
            float a01, a11, a10, a00, b01, b11, b10, b00;
            tld4_layer_f32(tex, max(iz, 0), lo(cxy), hi(cxy), a01, a11, a10, a00);
            tld4_layer_f32(tex, min(iz + 1, zmax), lo(cxy), hi(cxy), b01, b11, b10, b00);
            const f2 loA = mk2(a00, b00), hiA = mk2(a10, b10), loB = mk2(a01, b01), hiB = mk2(a11, b11);
            cA = ffma(wxx, fsub(hiA, loA), loA);
            cB = ffma(wxx, fsub(hiB, loB), loB);
        } else {
            uint32_t a01, a11, a10, a00, b01, b11, b10, b00;           // slice z (a) and z+1 (b); suffix = x,y offsets
            tld4_layer(tex, max(iz, 0), lo(cxy), hi(cxy), a01, a11, a10, a00);
            tld4_layer(tex, min(iz + 1, zmax), lo(cxy), hi(cxy), b01, b11, b10, b00);
            // pairs over (z, z+1): A = row y, B = row y+1
            const f2 loA = mk2(__uint_as_float(0x4B000000u | a00), __uint_as_float(0x4B000000u | b00));
            const f2 hiA = mk2(__uint_as_float(0x4B000000u | a10), __uint_as_float(0x4B000000u | b10));
            const f2 loB = mk2(__uint_as_float(0x4B000000u | a01), __uint_as_float(0x4B000000u | b01));
            const f2 hiB = mk2(__uint_as_float(0x4B000000u | a11), __uint_as_float(0x4B000000u | b11));
            cA = ffma(wxx, fsub(hiA, loA), fsub(loA, B2));
            cB = ffma(wxx, fsub(hiB, loB), fsub(loB, B2));
        }
        const f2 cy = ffma(wyy, fsub(cB, cA), cA);
        const float s = __fmaf_rn(wz, __fsub_rn(hi(cy), lo(cy)), lo(cy));
        float v;
        if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fc.fmin), fc.fmax), fc.fmin), fc.frange, fc.inv_frange);
        const float a = __fmul_rn(v, fc.alpha_scale);
        const float c = __fmul_rn(v, a);
        const float t = __fsub_rn(1.0f, A);
        const f2 ca_t = fmul(mk2(c, a), splat2(t));
        C = __fadd_rn(C, lo(ca_t));
        A = __fadd_rn(A, hi(ca_t));
        pxy = fadd(pxy, dxy);
        pz = __fadd_rn(pz, dz);
    }
    outC = C; outA = A;
}

struct TexArgs {
    cudaTextureObject_t tex;
    float* out;
    int local_rows;
    // linear, edge-replicated copy of the z-pair words for the LSU stage of the hybrid kernel:
    // (Nx+2) x (Ny+2) x (Nz+1) words, row pitch `zpitch`, slice stride `zslice` (words, < 2^31 in total)
    const void* zlin;
    int zpitch, zslice;
    const float* tf_lut;        // 256-entry opacity LUT (transfer-function form of the pipelined kernel)
};

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, bool FLOATTEX = false>
__global__ void __launch_bounds__(256, NOCAP ? 8 : 6)
march_texgather_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ TexArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (mapping each 4-lane quad to a 2x2 pixel block was measured: no change on K2, +-3 % on K0/K1)
    const int qx = lane & 7, qy = lane >> 3;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + qx;
    const int lrow = blockIdx.y * 8 + (warp >> 2) * 4 + qy;
    if (px >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const RaySetup r = setup_ray(fc, px, py);
    float C = 0.0f, A = 0.0f;
    if (r.hit) {
        float pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
            ds[i] = __fmul_rn(r.dir[i], fc.step);
        }
        march_ray_texgather<T, TCDIV, WIN, UNIT, NOCAP, FLOATTEX>(fc, args.tex, pos, ds, C, A);
    }
    const int orow = fc.compact ? lrow : py;
    reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
}

// ------------------------------------------------------------------------------------------
// z-pair texture-gather march ("texpair").  Third volume copy: a layered 2-D array of 32-bit
// (16-bit for 8-bit data) texels, layer L in [0, Nz], texel (x, y, L) =
//     v(x, y, max(L-1, 0))  |  v(x, y, min(L, Nz-1)) << bits
// so that the layer iz+1 (iz = floor(fz) in [-1, Nz-1]) carries BOTH z slices of the trilinear
// footprint with GL_CLAMP_TO_EDGE in z already applied.  ONE `tld4` per sample returns all
// eight texels; x/y clamp-to-edge is the texture unit's address mode.  The texel offset
// operand of tld4 (AOFFI, immediate {1,1}) moves the footprint from {i-1, i} to {i, i+1}, so
// the gather coordinate is (float(ix), float(iy)) -- the corner shared by the four texels,
// exact, half a texel from every footprint boundary -- and the `+1` add disappears.
// Per sample vs march_ray_texgather: -1 tld4, -1 coordinate vector, -2 layer clamps, -1 add;
// the TEX data pipe (76 % busy there) sees half the requests.  Costs 2x the source bytes in HBM.
__device__ __forceinline__ void tld4_pair(cudaTextureObject_t tex, int layer, float x, float y,
                                          uint32_t& t_x0y1, uint32_t& t_x1y1, uint32_t& t_x1y0, uint32_t& t_x0y0)
{
    asm volatile("tld4.r.a2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}], {1, 1};"
                 : "=r"(t_x0y1), "=r"(t_x1y1), "=r"(t_x1y0), "=r"(t_x0y0)
                 : "l"(tex), "r"(layer), "f"(x), "f"(y));
}

// biased floats 2^23 + v of the two halves of a z-pair texel: one PRMT each
template <typename T> __device__ __forceinline__ f2 unpack_zpair(uint32_t w);
template <> __device__ __forceinline__ f2 unpack_zpair<uint16_t>(uint32_t w)
{
    return mk2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)));
}
template <> __device__ __forceinline__ f2 unpack_zpair<uint8_t>(uint32_t w)
{
    return mk2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)));
}

// C/A carry the ray's state in and out and `iter` counts samples already taken, so the loop also
// finishes a ray whose partner in the two-ray kernel left the box first.
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP>
__device__ __forceinline__ void march_ray_texpair(const FrameConsts& fc, cudaTextureObject_t tex,
                                                  const float pos0[3], const float dstep[3], float& outC, float& outA,
                                                  int iter0 = 0)
{
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 hxy = mk2(fc.half_len[0], fc.half_len[1]);
    const float hz = fc.half_len[2];
    const f2 ixy = mk2(fc.inv_denom[0], fc.inv_denom[1]);
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const f2 mhalf = splat2(-0.5f), B2 = splat2(8388608.0f);
    float C = outC, A = outA;
    for (int iter = iter0; NOCAP || iter < 10000; ++iter) {
        const f2 qxy = fadd(pxy, hxy);
        const float qz = __fadd_rn(pz, hz);
        f2 txy;
        float tzq;
        if (UNIT) { txy = qxy; tzq = qz; }
        else if (TCDIV == DIV_RECIP_EXACT) { txy = fmul(qxy, ixy); tzq = __fmul_rn(qz, fc.inv_denom[2]); }
        else {
            const f2 q0 = fmul(qxy, ixy);
            const f2 r = ffma(mk2(-fc.denom[0], -fc.denom[1]), q0, qxy);
            txy = ffma(r, ixy, q0);
            tzq = div_by<DIV_MARKSTEIN>(qz, fc.denom[2], fc.inv_denom[2]);
        }
        const float tz = __fsub_rn(1.0f, tzq);
        const unsigned m = max(max(__float_as_uint(lo(txy)), __float_as_uint(hi(txy))), __float_as_uint(tz));
        if (m > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) break;           // :118
        const f2 fxy = ffma(txy, nxy, mhalf);
        const float fz = __fmaf_rn(tz, nz, -0.5f);
        const int ix = __float2int_rd(lo(fxy)), iy = __float2int_rd(hi(fxy)), iz = __float2int_rd(fz);
        const float flx = (float)ix, fly = (float)iy;
        uint32_t t01, t11, t10, t00;                                   // suffix = x,y offsets; halves = (z, z+1)
        tld4_pair(tex, iz + 1, flx, fly, t01, t11, t10, t00);
        const float wx = __fsub_rn(lo(fxy), flx), wy = __fsub_rn(hi(fxy), fly);
        const float wz = __fsub_rn(fz, (float)iz);
        const f2 wxx = splat2(wx), wyy = splat2(wy);
        const f2 loA = unpack_zpair<T>(t00), hiA = unpack_zpair<T>(t10);   // row y
        const f2 loB = unpack_zpair<T>(t01), hiB = unpack_zpair<T>(t11);   // row y+1
        const f2 cA = ffma(wxx, fsub(hiA, loA), fsub(loA, B2));
        const f2 cB = ffma(wxx, fsub(hiB, loB), fsub(loB, B2));
        const f2 cy = ffma(wyy, fsub(cB, cA), cA);
        const float s = __fmaf_rn(wz, __fsub_rn(hi(cy), lo(cy)), lo(cy));
        float v;
        if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fc.fmin), fc.fmax), fc.fmin), fc.frange, fc.inv_frange);
        const float a = __fmul_rn(v, fc.alpha_scale);
        const float c = __fmul_rn(v, a);
        const float t = __fsub_rn(1.0f, A);
        const f2 ca_t = fmul(mk2(c, a), splat2(t));
        C = __fadd_rn(C, lo(ca_t));
        A = __fadd_rn(A, hi(ca_t));
        pxy = fadd(pxy, dxy);
        pz = __fadd_rn(pz, dz);
    }
    outC = C; outA = A;
}

// Software-pipelined form of march_ray_texpair.  The loop above keeps ONE tld4 in flight per
// warp; 64 warps x 16 writeback cycles per tld4 then cover only ~73 % of the texture unit's
// return path while every warp waits a full queue length for its own fetch (ncu:
// l1tex__tex_writeback_active 73 %, long_scoreboard).  The gather coordinates of sample i+1
// depend only on `pos`, never on fetched data, so its tld4 is issued BEFORE sample i is
// unpacked, interpolated and composited: two fetches in flight per warp.  Scheduling only --
// the operation sequence per sample, and hence every bit of the result, is unchanged.  The
// look-ahead fetch is skipped when sample i+1 lies outside the box.
//
// FA / FB select the fetch path of the two ping-pong stages (even / odd samples): FETCH_TEX is the tld4
// above; FETCH_LSU reads the same four z-pair words with 4 ld.global.nc from a linear edge-replicated
// copy.  The texture unit returns 32 B/clk/SM (measured: 16 writeback cycles per 32-bit tld4 warp
// instruction) and is ~75 % busy when it serves every sample; alternating the two paths halves its
// load and gives the otherwise idle LSU/L1 data pipe the other half.
// lab only (upper bound for run-time specialisation): headline-frame constants as immediates
#if defined(VR_LAB_IMM) && VR_LAB_IMM >= 1
#define VR_KG(rt, imm) (imm)
#else
#define VR_KG(rt, imm) (rt)
#endif
#if defined(VR_LAB_IMM) && VR_LAB_IMM >= 2
#define VR_KW(rt, imm) (imm)
#else
#define VR_KW(rt, imm) (rt)
#endif
enum { FETCH_TEX = 0, FETCH_LSU = 1 };
// MODE: MODE_DVR = rayMarchVolume (:104-139); MODE_TF = the same with the transfer-function extension
// (SURVEY 8a-7): src.a = lut[floor(v*255 + 0.5)], rgb stays v; MODE_MIP = MIP (:141-173): dest = max over
// the samples of v*alpha_scale, with the inherited `dest.a >= 0.95` exit of :156; MODE_DVR_TOP / _BOTTOM =
// DVR with the view_top / view_bottom tex-coord swizzle of cartesianToTextureCoord (:183-190; the host
// builds the bounding box from the .xzy-swizzled dims and spacing, :68-78).
enum { MODE_DVR = 0, MODE_TF = 1, MODE_MIP = 2, MODE_DVR_TOP = 3, MODE_DVR_BOTTOM = 4 };
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, int FA = FETCH_TEX, int FB = FETCH_TEX, int MODE = MODE_DVR>
__device__ __forceinline__ void march_ray_texpair_pipe(const FrameConsts& fc, const TexArgs& args,
                                                       const float pos0[3], const float dstep[3], float& outC, float& outA,
                                                       int iter0 = 0)
{
    const cudaTextureObject_t tex = args.tex;
    typedef typename PairWord<T>::type ZW;
    const int zpitch = args.zpitch, zslice = args.zslice;
    const ZW* __restrict__ zbase = static_cast<const ZW*>(args.zlin) + ((size_t)zslice + zpitch + 1);   // (ix, iy, iz) = (-1, -1, -1) lands on word 0
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 hxy = mk2(VR_KG(fc.half_len[0], 0.5f), VR_KG(fc.half_len[1], 0.5f));
    const float hz = VR_KG(fc.half_len[2], 0.5f);
    const f2 ixy = mk2(fc.inv_denom[0], fc.inv_denom[1]);
    const f2 nxy = mk2(VR_KG(fc.dimf[0], 1024.0f), VR_KG(fc.dimf[1], 1024.0f));
    const float nz = VR_KG(fc.dimf[2], 1024.0f);
    const f2 mhalf = splat2(-0.5f), B2 = splat2(8388608.0f);
    float C = outC, A = outA;

    // tex-coord of the sample at (pxy, pz): cartesianToTextureCoord :175-192; returns the :118 range key
    auto tex_coord_key = [&](f2& txy, float& tz) -> unsigned {
        const f2 qxy = fadd(pxy, hxy);
        const float qz = __fadd_rn(pz, hz);
        float tzq;
        if (UNIT) { txy = qxy; tzq = qz; }
        else if (TCDIV == DIV_RECIP_EXACT) { txy = fmul(qxy, ixy); tzq = __fmul_rn(qz, fc.inv_denom[2]); }
        else {
            const f2 q0 = fmul(qxy, ixy);
            const f2 r = ffma(mk2(-fc.denom[0], -fc.denom[1]), q0, qxy);
            txy = ffma(r, ixy, q0);
            tzq = div_by<DIV_MARKSTEIN>(qz, fc.denom[2], fc.inv_denom[2]);
        }
        tz = __fsub_rn(1.0f, tzq);                                      // :180
        if (MODE == MODE_DVR_TOP) {                                     // :183-186  (x, 1 - z, y), z already flipped
            const float ty = __fsub_rn(1.0f, tz), qy = hi(txy);
            txy = mk2(lo(txy), ty); tz = qy;
        } else if (MODE == MODE_DVR_BOTTOM) {                           // :187-190  (x, z, 1 - y)
            const float ty = tz, qy = hi(txy);
            txy = mk2(lo(txy), ty); tz = __fsub_rn(1.0f, qy);
        }
        return max(max(__float_as_uint(lo(txy)), __float_as_uint(hi(txy))), __float_as_uint(tz));
    };

    struct Fetched { uint32_t t01, t11, t10, t00; float wx, wy, wz; };
    const unsigned last_layer = VR_KG((unsigned)fc.dim[2], 1024u);

    // texel coordinates, weights and the gather of the sample whose tex-coord is (txy, tz).  The layer is
    // clamped because a look-ahead sample may lie one step outside the box (its texels are never used).
    auto fetch = [&](int path, f2 txy, float tz, bool inside, Fetched& f) {
        const f2 fxy = ffma(txy, nxy, mhalf);
        const float fz = __fmaf_rn(tz, nz, -0.5f);
        const int ix = __float2int_rd(lo(fxy)), iy = __float2int_rd(hi(fxy)), iz = __float2int_rd(fz);
        const float flx = (float)ix, fly = (float)iy;
        if (path == FETCH_TEX) {
            tld4_pair(tex, (int)min((unsigned)(iz + 1), last_layer), flx, fly, f.t01, f.t11, f.t10, f.t00);
        } else {
            // an outside look-ahead sample reads word 0 instead (its texels are never used)
            const int e = inside ? iz * zslice + (iy * zpitch + ix) : -(zslice + zpitch + 1);
            const ZW* __restrict__ p0 = zbase + e;
            const ZW* __restrict__ p1 = zbase + (e + zpitch);
            f.t00 = __ldg(p0); f.t10 = __ldg(p0 + 1); f.t01 = __ldg(p1); f.t11 = __ldg(p1 + 1);
        }
        f.wx = __fsub_rn(lo(fxy), flx); f.wy = __fsub_rn(hi(fxy), fly); f.wz = __fsub_rn(fz, (float)iz);
    };
    // interpolate, window, composite one fetched sample (:121-132)
    auto consume = [&](const Fetched& f) {
        const f2 wxx = splat2(f.wx), wyy = splat2(f.wy);
        const f2 loA = unpack_zpair<T>(f.t00), hiA = unpack_zpair<T>(f.t10);   // row y
        const f2 loB = unpack_zpair<T>(f.t01), hiB = unpack_zpair<T>(f.t11);   // row y+1
        const f2 cA = ffma(wxx, fsub(hiA, loA), fsub(loA, B2));
        const f2 cB = ffma(wxx, fsub(hiB, loB), fsub(loB, B2));
        const f2 cy = ffma(wyy, fsub(cB, cA), cA);
        const float s = __fmaf_rn(f.wz, __fsub_rn(hi(cy), lo(cy)), lo(cy));
        float v;
        if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, VR_KW(fc.frange, 4095.0f), VR_KW(fc.inv_frange, 1.0f / 4095.0f));
        else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fc.fmin), fc.fmax), fc.fmin), fc.frange, fc.inv_frange);
        if (MODE == MODE_MIP) {
            const float m = __fmul_rn(v, VR_KW(fc.alpha_scale, 0.02f));            // :163
            if (A < m) { C = m; A = m; }                                           // :164-167
            return;
        }
        float src_a = v;
        if (MODE == MODE_TF) {
            // v in [0,1] on this path (ordered window), so the index is in [0,255]; the unsigned min only guards the load
            const unsigned iso = min((unsigned)__float2int_rd(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f)), 255u);
            src_a = __ldg(args.tf_lut + iso);
        }
        const float a = __fmul_rn(src_a, VR_KW(fc.alpha_scale, 0.02f));
        const float c = __fmul_rn(v, a);
        const float t = __fsub_rn(1.0f, A);
        const f2 ca_t = fmul(mk2(c, a), splat2(t));
        C = __fadd_rn(C, lo(ca_t));
        A = __fadd_rn(A, hi(ca_t));
    };
    // one pipeline stage: look ahead to sample i+1 (advance :136, range key :118, gather), then finish
    // sample i.  Returns false when sample i+1 must not be taken.
    auto stage = [&](int path_n, const Fetched& cur, Fetched& nxt) -> bool {
        pxy = fadd(pxy, dxy);
        pz = __fadd_rn(pz, dz);
        f2 txy; float tz;
        const bool inside_n = tex_coord_key(txy, tz) <= 0x3F800000u;
        fetch(path_n, txy, tz, inside_n, nxt);
        consume(cur);
        return inside_n && __float_as_uint(A) < 0x3F733333u;
    };

    {
        f2 txy; float tz;
        if (tex_coord_key(txy, tz) > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return;   // :118, first sample
        Fetched fa, fb;                                                // ping-pong: no register rotation
        fetch(FA, txy, tz, true, fa);
        for (int iter = iter0; NOCAP || iter < 10000; iter += 2) {
            if (!stage(FB, fa, fb)) break;
            if (!NOCAP && iter + 1 >= 10000) break;
            if (!stage(FA, fb, fa)) break;
        }
    }
    outC = C; outA = A;
}

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, int MINB, int FA = FETCH_TEX, int FB = FETCH_TEX, int MODE = MODE_DVR>
__global__ void __launch_bounds__(256, MINB)
march_texpair_pipe_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ TexArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int lrow = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const RaySetup r = setup_ray(fc, px, py);
    float C = 0.0f, A = 0.0f;
    if (r.hit) {
        float pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
            ds[i] = __fmul_rn(r.dir[i], fc.step);
        }
        march_ray_texpair_pipe<T, TCDIV, WIN, UNIT, NOCAP, FA, FB, MODE>(fc, args, pos, ds, C, A);
    }
    const int orow = fc.compact ? lrow : py;
    reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
}

// ------------------------------------------------------------------------------------------
// Nearest-filter march (what an integer texture with GL_LINEAR does on NVIDIA drivers, i.e. the
// reference's de-facto output): one texel per sample, fetched with an integer-coordinate texel load
// (`tex.a2d ... .s32` -> SASS TLD.LZ, one component returned) from the source-type layered array.
// i = clamp(floor(u * N), 0, N-1) per axis (VolumeRenderer.cs:121 + GL_CLAMP_TO_EDGE): the clamp is an
// unsigned min, which also makes the look-ahead fetch of the software pipeline safe when sample i+1
// lies outside the box.  Same ping-pong pipeline as march_ray_texpair_pipe (two loads in flight per warp).
__device__ __forceinline__ uint32_t tld_layer(cudaTextureObject_t tex, unsigned layer, unsigned x, unsigned y)
{
    uint32_t r, g, b, a;
    asm volatile("tex.a2d.v4.u32.s32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                 : "=r"(r), "=r"(g), "=r"(b), "=r"(a) : "l"(tex), "r"(layer), "r"(x), "r"(y));
    return r;
}

template <int TCDIV, int WIN, bool UNIT, bool NOCAP>
__device__ __forceinline__ void march_ray_nearest_tex(const FrameConsts& fc, cudaTextureObject_t tex,
                                                      const float pos0[3], const float dstep[3], float& outC, float& outA)
{
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 hxy = mk2(fc.half_len[0], fc.half_len[1]);
    const float hz = fc.half_len[2];
    const f2 ixy = mk2(fc.inv_denom[0], fc.inv_denom[1]);
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const unsigned mx = (unsigned)fc.dim[0] - 1u, my = (unsigned)fc.dim[1] - 1u, mz = (unsigned)fc.dim[2] - 1u;
    float C = outC, A = outA;

    auto tex_coord_key = [&](f2& txy, float& tz) -> unsigned {       // cartesianToTextureCoord :175-192
        const f2 qxy = fadd(pxy, hxy);
        const float qz = __fadd_rn(pz, hz);
        float tzq;
        if (UNIT) { txy = qxy; tzq = qz; }
        else if (TCDIV == DIV_RECIP_EXACT) { txy = fmul(qxy, ixy); tzq = __fmul_rn(qz, fc.inv_denom[2]); }
        else {
            const f2 q0 = fmul(qxy, ixy);
            const f2 r = ffma(mk2(-fc.denom[0], -fc.denom[1]), q0, qxy);
            txy = ffma(r, ixy, q0);
            tzq = div_by<DIV_MARKSTEIN>(qz, fc.denom[2], fc.inv_denom[2]);
        }
        tz = __fsub_rn(1.0f, tzq);
        return max(max(__float_as_uint(lo(txy)), __float_as_uint(hi(txy))), __float_as_uint(tz));
    };
    auto fetch = [&](f2 txy, float tz) -> uint32_t {
        const f2 uxy = fmul(txy, nxy);
        const float uz = __fmul_rn(tz, nz);
        const unsigned ix = min((unsigned)__float2int_rd(lo(uxy)), mx);
        const unsigned iy = min((unsigned)__float2int_rd(hi(uxy)), my);
        const unsigned iz = min((unsigned)__float2int_rd(uz), mz);
        return tld_layer(tex, iz, ix, iy);
    };
    auto consume = [&](uint32_t texel) {                               // :121-132
        const float s = __fsub_rn(__uint_as_float(0x4B000000u | texel), 8388608.0f);
        float v;
        if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fc.fmin), fc.fmax), fc.fmin), fc.frange, fc.inv_frange);
        const float a = __fmul_rn(v, fc.alpha_scale);
        const float c = __fmul_rn(v, a);
        const float t = __fsub_rn(1.0f, A);
        const f2 ca_t = fmul(mk2(c, a), splat2(t));
        C = __fadd_rn(C, lo(ca_t));
        A = __fadd_rn(A, hi(ca_t));
    };
    auto stage = [&](uint32_t cur, uint32_t& nxt) -> bool {
        pxy = fadd(pxy, dxy);                                          // :136
        pz = __fadd_rn(pz, dz);
        f2 txy; float tz;
        const bool inside_n = tex_coord_key(txy, tz) <= 0x3F800000u;   // :118 of sample i+1
        nxt = fetch(txy, tz);
        consume(cur);
        return inside_n && __float_as_uint(A) < 0x3F733333u;
    };

    f2 txy; float tz;
    if (tex_coord_key(txy, tz) > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return;   // :118, first sample
    uint32_t ta = fetch(txy, tz), tb = 0;
    for (int iter = 0; NOCAP || iter < 10000; iter += 2) {
        if (!stage(ta, tb)) break;
        if (!NOCAP && iter + 1 >= 10000) break;
        if (!stage(tb, ta)) break;
    }
    outC = C; outA = A;
}

template <int TCDIV, int WIN, bool UNIT, bool NOCAP>
__global__ void __launch_bounds__(256, NOCAP ? 8 : 6)
march_nearest_tex_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ TexArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int lrow = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const RaySetup r = setup_ray(fc, px, py);
    float C = 0.0f, A = 0.0f;
    if (r.hit) {
        float pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
            ds[i] = __fmul_rn(r.dir[i], fc.step);
        }
        march_ray_nearest_tex<TCDIV, WIN, UNIT, NOCAP>(fc, args.tex, pos, ds, C, A);
    }
    const int orow = fc.compact ? lrow : py;
    reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
}

// Two rays (horizontally adjacent pixels) per thread over the same z-pair array: every IEEE
// operation of the sample -- position, tex-coord, texel coordinate, weights, the seven lerps,
// window, compositing products -- is issued ONCE as a packed f32x2 instruction for both rays
// (lane .x = even pixel, .y = odd pixel); per thread and iteration 2 tld4 fetch 16 texels.  Sums
// fed by an unfused product stay scalar (ptxas would contract them).  Returns the `done` bits;
// the survivor of a pair is finished by march_ray_texpair.
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP>
__device__ __forceinline__ unsigned march_rays2_texpair(const FrameConsts& fc, cudaTextureObject_t tex,
                                                        f2 pos[3], const f2 ds[3], f2& C, f2& A, int& iter)
{
    const f2 hx = splat2(fc.half_len[0]), hy = splat2(fc.half_len[1]), hz = splat2(fc.half_len[2]);
    const f2 nx = splat2(fc.dimf[0]), ny = splat2(fc.dimf[1]), nz = splat2(fc.dimf[2]);
    const f2 mhalf = splat2(-0.5f), B2 = splat2(8388608.0f), one2 = splat2(1.0f);
    const f2 alpha = splat2(fc.alpha_scale), vfmin = splat2(fc.fmin);
    unsigned done = 0;
    for (; NOCAP || iter < 10000; ++iter) {
        const f2 qx = fadd(pos[0], hx), qy = fadd(pos[1], hy), qz = fadd(pos[2], hz);
        f2 tx, ty, tz;
        if (UNIT) { tx = qx; ty = qy; tz = fsub(one2, qz); }
        else {
            tx = div_by_l<TCDIV>(qx, fc.denom[0], fc.inv_denom[0]);
            ty = div_by_l<TCDIV>(qy, fc.denom[1], fc.inv_denom[1]);
            tz = fsub_after_mul(one2, div_by_l<TCDIV>(qz, fc.denom[2], fc.inv_denom[2]));
        }
        // :118 for both rays at once (bit patterns; no -0 / NaN on this path); which ray it was
        // is worked out after the loop
        const unsigned m0 = max(max(__float_as_uint(lo(tx)), __float_as_uint(lo(ty))), __float_as_uint(lo(tz)));
        const unsigned m = max(max(m0, __float_as_uint(hi(tx))), max(__float_as_uint(hi(ty)), __float_as_uint(hi(tz))));
        if (m > 0x3F800000u || max(__float_as_uint(lo(A)), __float_as_uint(hi(A))) >= 0x3F733333u) {
            const unsigned m1 = max(max(__float_as_uint(hi(tx)), __float_as_uint(hi(ty))), __float_as_uint(hi(tz)));
            if (m0 > 0x3F800000u || __float_as_uint(lo(A)) >= 0x3F733333u) done |= 1u;
            if (m1 > 0x3F800000u || __float_as_uint(hi(A)) >= 0x3F733333u) done |= 2u;
            break;
        }
        const f2 fx = ffma(tx, nx, mhalf), fy = ffma(ty, ny, mhalf), fz = ffma(tz, nz, mhalf);
        const int ix0 = __float2int_rd(lo(fx)), iy0 = __float2int_rd(lo(fy)), iz0 = __float2int_rd(lo(fz));
        const int ix1 = __float2int_rd(hi(fx)), iy1 = __float2int_rd(hi(fy)), iz1 = __float2int_rd(hi(fz));
        const float flx0 = (float)ix0, fly0 = (float)iy0, flx1 = (float)ix1, fly1 = (float)iy1;
        uint32_t a01, a11, a10, a00, b01, b11, b10, b00;               // ray 0 (a), ray 1 (b); suffix = x,y offsets
        tld4_pair(tex, iz0 + 1, flx0, fly0, a01, a11, a10, a00);
        tld4_pair(tex, iz1 + 1, flx1, fly1, b01, b11, b10, b00);
        const f2 wx = fsub(fx, mk2(flx0, flx1)), wy = fsub(fy, mk2(fly0, fly1));
        const f2 wz = fsub(fz, mk2((float)iz0, (float)iz1));
        // biased floats 2^23 + v, paired across the two rays: L = slice z, H = slice z+1
        const f2 za00 = unpack_zpair<T>(a00), za10 = unpack_zpair<T>(a10), za01 = unpack_zpair<T>(a01), za11 = unpack_zpair<T>(a11);
        const f2 zb00 = unpack_zpair<T>(b00), zb10 = unpack_zpair<T>(b10), zb01 = unpack_zpair<T>(b01), zb11 = unpack_zpair<T>(b11);
        const f2 L00 = mk2(lo(za00), lo(zb00)), H00 = mk2(hi(za00), hi(zb00));
        const f2 L10 = mk2(lo(za10), lo(zb10)), H10 = mk2(hi(za10), hi(zb10));
        const f2 L01 = mk2(lo(za01), lo(zb01)), H01 = mk2(hi(za01), hi(zb01));
        const f2 L11 = mk2(lo(za11), lo(zb11)), H11 = mk2(hi(za11), hi(zb11));
        const f2 c00 = ffma(wx, fsub(L10, L00), fsub(L00, B2));
        const f2 c10 = ffma(wx, fsub(L11, L01), fsub(L01, B2));
        const f2 c01 = ffma(wx, fsub(H10, H00), fsub(H00, B2));
        const f2 c11 = ffma(wx, fsub(H11, H01), fsub(H01, B2));
        const f2 c0 = ffma(wy, fsub(c10, c00), c00);
        const f2 c1 = ffma(wy, fsub(c11, c01), c01);
        const f2 s = ffma(wz, fsub(c1, c0), c0);
        f2 v;                                                           // :122-124
        if (WIN == WIN_COVERS0) v = div_by_l<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
        else {
            const f2 cl = mk2(fminf(fmaxf(lo(s), fc.fmin), fc.fmax), fminf(fmaxf(hi(s), fc.fmin), fc.fmax));
            v = div_by_l<DIV_MARKSTEIN>(fsub(cl, vfmin), fc.frange, fc.inv_frange);
        }
        const f2 a = fmul(v, alpha);                                    // :130-132
        const f2 c = fmul(v, a);
        const f2 t = fsub(one2, A);
        C = fadd_after_mul(C, fmul(c, t));
        A = fadd_after_mul(A, fmul(a, t));
        pos[0] = fadd(pos[0], ds[0]);                                   // :136
        pos[1] = fadd(pos[1], ds[1]);
        pos[2] = fadd(pos[2], ds[2]);
    }
    return done;
}

// CTA = 256 threads = 64 x 8 pixels; a warp covers 16 x 4 pixels (8 x 4 threads, 2 pixels each).
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, int MINB>
__global__ void __launch_bounds__(256, MINB)
march_texpair2_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ TexArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px0 = (blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7)) * 2;
    const int lrow = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px0 >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const bool have1 = (px0 + 1) < fc.W;
    const RaySetup r0 = setup_ray(fc, px0, py);
    const RaySetup r1 = setup_ray(fc, have1 ? px0 + 1 : px0, py);
    float p0[3], p1[3], d0[3], d1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        p0[i] = __fadd_rn(__fadd_rn(r0.org[i], __fmul_rn(r0.dir[i], r0.t_min)), __fmul_rn(r0.dir[i], 0.000001f));
        p1[i] = __fadd_rn(__fadd_rn(r1.org[i], __fmul_rn(r1.dir[i], r1.t_min)), __fmul_rn(r1.dir[i], 0.000001f));
        d0[i] = __fmul_rn(r0.dir[i], fc.step);
        d1[i] = __fmul_rn(r1.dir[i], fc.step);
    }
    float C0 = 0.f, A0 = 0.f, C1 = 0.f, A1 = 0.f;
    unsigned alive = (r0.hit ? 1u : 0u) | ((r1.hit && have1) ? 2u : 0u);
    int iter = 0;
    if (alive == 3u) {
        f2 pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { pos[i] = mk2(p0[i], p1[i]); ds[i] = mk2(d0[i], d1[i]); }
        f2 C = mk2(0.f, 0.f), A = mk2(0.f, 0.f);
        const unsigned done = march_rays2_texpair<T, TCDIV, WIN, UNIT, NOCAP>(fc, args.tex, pos, ds, C, A, iter);
        C0 = lo(C); C1 = hi(C); A0 = lo(A); A1 = hi(A);
#pragma unroll
        for (int i = 0; i < 3; ++i) { p0[i] = lo(pos[i]); p1[i] = hi(pos[i]); }
        alive &= ~done;
        if (!NOCAP && iter >= 10000) alive = 0;
    }
    if (alive) {   // finish whichever ray is still marching
        const bool second = (alive & 2u) != 0;
        float pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { pos[i] = second ? p1[i] : p0[i]; ds[i] = second ? d1[i] : d0[i]; }
        float C = second ? C1 : C0, A = second ? A1 : A0;
        march_ray_texpair<T, TCDIV, WIN, UNIT, NOCAP>(fc, args.tex, pos, ds, C, A, iter);
        if (second) { C1 = C; A1 = A; } else { C0 = C; A0 = A; }
    }
    const int orow = fc.compact ? lrow : py;
    float4* out = reinterpret_cast<float4*>(args.out) + (size_t)orow * fc.W;
    out[px0] = make_float4(C0, C0, C0, A0);
    if (have1) out[px0 + 1] = make_float4(C1, C1, C1, A1);
}

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP>
__global__ void __launch_bounds__(256, NOCAP ? 8 : 6)
march_texpair_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ TexArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int lrow = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const RaySetup r = setup_ray(fc, px, py);
    float C = 0.0f, A = 0.0f;
    if (r.hit) {
        float pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
            ds[i] = __fmul_rn(r.dir[i], fc.step);
        }
        march_ray_texpair<T, TCDIV, WIN, UNIT, NOCAP>(fc, args.tex, pos, ds, C, A);
    }
    const int orow = fc.compact ? lrow : py;
    reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
}

// launch bounds measured on the headline frame: the capless loop fits 32 registers (8 CTAs/SM,
// 3.50 ms); with the iteration counter 32 registers spill (4.31 ms) and 40 are best (3.78 ms)
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP>
__global__ void __launch_bounds__(256, NOCAP ? 8 : 6)
march_packed_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ FastArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int lrow = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (px >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const RaySetup r = setup_ray(fc, px, py);
    float C = 0.0f, A = 0.0f;
    if (r.hit) {
        float pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
            ds[i] = __fmul_rn(r.dir[i], fc.step);
        }
        march_ray_packed<T, TCDIV, WIN, UNIT, NOCAP>(fc, static_cast<const T*>(args.vol), args.pitch, args.slice_lo, pos, ds, C, A);
    }
    const int orow = fc.compact ? lrow : py;
    reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
}

constexpr int FAST_THREADS = 256;

// RAYS = 2: a thread owns pixels (2i, 2i+1) of a row; a warp covers 16x4 pixels, a CTA 64x8.
// RAYS = 1: a warp covers 8x4 pixels, a CTA 32x8 (same as the baseline kernel).
// (forcing 32 registers/thread with __launch_bounds__(256, 8) spills and is slower: 4.24 ms vs 3.97 ms)
template <typename T, int FILTER, int TCDIV, int WIN, int FM, int RAYS, bool PAIRS, bool PIPE>
__device__ __forceinline__ void fast_tile(const FrameConsts& fc, const FastArgs& args, int bx, int by, int warp)
{
    const int lane = threadIdx.x & 31;
    const int tx = bx * 32 + (warp & 3) * 8 + (lane & 7);       // thread column
    const int lrow = by * 8 + (warp >> 2) * 4 + (lane >> 3);
    const int px0 = tx * RAYS;
    if (px0 >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;
    const T* __restrict__ vol = static_cast<const T*>(args.vol);
    const uint32_t pitch = args.pitch, slice = args.slice_lo;
    const int orow = fc.compact ? lrow : py;
    float4* out = reinterpret_cast<float4*>(args.out) + (size_t)orow * fc.W;

    if (RAYS == 1) {
        const RaySetup r = setup_ray(fc, px0, py);
        f1 C{0.0f}, A{0.0f};
        if (r.hit) {
            f1 pos[3], ds[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pos[i].v = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
                ds[i].v = __fmul_rn(r.dir[i], fc.step);
            }
            int iter = 0;
            if (PIPE && FILTER == VR_FILTER_TRILINEAR) march_lanes_pipe<T, TCDIV, WIN, FM, f1, PAIRS>(fc, vol, pitch, slice, pos, ds, C, A, iter);
            else march_lanes<T, FILTER, TCDIV, WIN, FM, f1, PAIRS>(fc, vol, pitch, slice, pos, ds, C, A, iter);
        }
        out[px0] = make_float4(C.v, C.v, C.v, A.v);
        return;
    }

    // ---- two rays ----
    const bool have1 = (px0 + 1) < fc.W;
    const RaySetup r0 = setup_ray(fc, px0, py);
    const RaySetup r1 = setup_ray(fc, have1 ? px0 + 1 : px0, py);
    float p0[3], p1[3], d0[3], d1[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        p0[i] = __fadd_rn(__fadd_rn(r0.org[i], __fmul_rn(r0.dir[i], r0.t_min)), __fmul_rn(r0.dir[i], 0.000001f));
        p1[i] = __fadd_rn(__fadd_rn(r1.org[i], __fmul_rn(r1.dir[i], r1.t_min)), __fmul_rn(r1.dir[i], 0.000001f));
        d0[i] = __fmul_rn(r0.dir[i], fc.step);
        d1[i] = __fmul_rn(r1.dir[i], fc.step);
    }
    float C0 = 0.f, A0 = 0.f, C1 = 0.f, A1 = 0.f;
    unsigned alive = (r0.hit ? 1u : 0u) | ((r1.hit && have1) ? 2u : 0u);
    int iter = 0;
    if (alive == 3u) {
        f2 pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { pos[i] = mk2(p0[i], p1[i]); ds[i] = mk2(d0[i], d1[i]); }
        f2 C = mk2(0.f, 0.f), A = mk2(0.f, 0.f);
        const unsigned done = (PIPE && FILTER == VR_FILTER_TRILINEAR)
            ? march_lanes_pipe<T, TCDIV, WIN, FM, f2, PAIRS>(fc, vol, pitch, slice, pos, ds, C, A, iter)
            : march_lanes<T, FILTER, TCDIV, WIN, FM, f2, PAIRS>(fc, vol, pitch, slice, pos, ds, C, A, iter);
        C0 = lo(C); C1 = hi(C); A0 = lo(A); A1 = hi(A);
#pragma unroll
        for (int i = 0; i < 3; ++i) { p0[i] = lo(pos[i]); p1[i] = hi(pos[i]); }
        alive &= ~done;
        if (iter >= 10000) alive = 0;
    }
    if (alive) {   // finish whichever ray is still marching (both rays share `iter` so far)
        const bool second = (alive & 2u) != 0;
        f1 pos[3], ds[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { pos[i].v = second ? p1[i] : p0[i]; ds[i].v = second ? d1[i] : d0[i]; }
        f1 C{second ? C1 : C0}, A{second ? A1 : A0};
        if (PIPE && FILTER == VR_FILTER_TRILINEAR) march_lanes_pipe<T, TCDIV, WIN, FM, f1, PAIRS>(fc, vol, pitch, slice, pos, ds, C, A, iter);
        else march_lanes<T, FILTER, TCDIV, WIN, FM, f1, PAIRS>(fc, vol, pitch, slice, pos, ds, C, A, iter);
        if (second) { C1 = C.v; A1 = A.v; } else { C0 = C.v; A0 = A.v; }
    }
    out[px0] = make_float4(C0, C0, C0, A0);
    if (have1) out[px0 + 1] = make_float4(C1, C1, C1, A1);
}

// One CTA per 32x8 (64x8 with two rays per thread) pixel tile.  Measured alternatives that did
// not pay (tools/marchlab, profiles/r01/experiments.md): persistent CTAs pulling tiles from an
// atomic counter (4.46 ms vs 3.99 ms), persistent WARPS pulling 8x4 patches (4.10 ms; 0.85 vs
// 0.89 ms on a 1/8 partition), long-rays-first tile order (no change).
template <typename T, int FILTER, int TCDIV, int WIN, int FM, int RAYS, bool PAIRS = false, bool PIPE = false>
__global__ void __launch_bounds__(FAST_THREADS)
march_fast_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ FastArgs args)
{
    fast_tile<T, FILTER, TCDIV, WIN, FM, RAYS, PAIRS, PIPE>(fc, args, blockIdx.x, blockIdx.y, threadIdx.x >> 5);
}

}  // namespace vr
