// kernel_march.cuh -- the optimised march kernels of the product (everything VR_KERNEL_AUTO runs
// except degenerate frames, which take the generic loop of kernel_direct.cuh):
//
//   march_texpair_kernel   trilinear filter.  One thread per ray; ONE texture gather (tld4) per
//                          sample from the z-pair array returns the eight texels of the footprint;
//                          software pipelined DEPTH deep; interpolation / window / compositing in
//                          fp32 ALU with f32x2 packing inside the ray.
//   march_nearest_kernel   nearest filter (what an integer texture with GL_LINEAR does, i.e. the
//                          reference's de-facto output): one integer-coordinate texel load (TLD)
//                          per sample from the source-type layered array, same pipeline.
//
// Both reproduce VolumeRenderer.cs:104-192 operation by operation (see march_device.cuh for the
// contract) -- results are bit-identical to the oracle -- in these FORMS of the same loop:
//   FORM_DVR      rayMarchVolume (:104-139), default view
//   FORM_TF       + transfer-function extension: src.a = lut[floor(v*255 + 0.5)]
//   FORM_MIP      MIP (:141-173), default view
//   FORM_GENERAL  run-time switches for view_top / view_bottom (:183-190), TF, MIP in any
//                 combination, and a host-finalised opacity LUT (TF + opacity correction)
//   FORM_GENERAL_OC  the same plus per-sample opacity correction a' = 1-(1-a)^step_scale
// and, orthogonally, SKIP: result-identical empty-space skipping.  A bit map (staged in shared
// memory by every CTA) marks the cells (2^s voxels cubed, s chosen at upload) in which every voxel
// a sample can touch is <= min_val; such a sample has v = 0 exactly (clamp, subtract min_val), adds
// exactly 0 to both accumulators (alpha_scale >= 0, lut[0]*alpha == 0) and never wins a MIP
// comparison, so neither its texels are fetched nor its arithmetic issued.  The map is consulted only
// at checkpoints (every few samples, at a warp-uniform point of the loop), not per sample.  When the front of the
// pipeline meets an empty cell it LEAPS: from the sample's exact index-space coordinate and the
// ray's per-step increment it bounds (conservatively, 0.05 voxel inside the cell and the box) how
// many further samples are certain to stay in that cell, then performs exactly that many position
// updates `pos += dir*step` (:136) -- the same rounded additions the reference performs, so the
// sample that follows is bit for bit the one the reference takes -- and re-examines the landing
// sample in full (box test :118, cell test).
#pragma once

#include "f32x2.cuh"
#include "march_device.cuh"

namespace vr {

enum WinMode : int {
    WIN_CLAMP = 0,     // clamp to [min,max], subtract min, divide by range (any ordered window)
    WIN_COVERS0 = 1    // min == 0 and every voxel value <= max: clamp and subtraction are no-ops
};

enum MarchForm : int { FORM_DVR = 0, FORM_TF = 1, FORM_MIP = 2, FORM_GENERAL = 3, FORM_GENERAL_OC = 4 };

struct MarchArgs {
    cudaTextureObject_t tex;     // z-pair array (trilinear) / source-type layered array (nearest)
    float* out;
    int local_rows;
    const float* tf_lut;         // 256-entry opacity LUT
    int lut_final;               // the LUT already holds the FINAL opacity (alpha-scaled and opacity-corrected on the host)
    // empty-cell bit map (SKIP forms): bit (cz*cell_ny + cy)*cell_nx + cx, c = (base voxel index + 1) >> cell_shift
    const uint32_t* cell_bits;
    int cell_words, cell_shift, cell_nx, cell_nxy;
    int skip_check_mask;         // checkpoints every (mask + 1)-th pass of the unrolled loop (mask + 1 a power of two)
    // launch order of the CTA tiles (small grids only; nullptr = row-major): entry b = tile (x | y << 16) taken by the
    // b-th CTA the hardware starts.  The host sorts the tiles by the estimated length of their rays, longest first, so
    // that what runs on the draining machine at the end of the grid are the short rays (LPT scheduling)
    const uint32_t* cta_order;
    // fused multi-GPU hand-off: the last CTA of the grid to finish publishes this rank's arrival in the
    // frame owner's barrier word (peer memory), replacing a one-thread kernel per frame
    unsigned int* done_counter;
    unsigned int* peer_arrive;
    unsigned int grid_ctas;
};

// ---- texture instructions ---------------------------------------------------------------------------
// z-pair array: layered 2-D array of 32-bit (16-bit for 8-bit data) texels, layer L in [0, Nz], texel
// (x, y, L) = v(x, y, max(L-1, 0)) | v(x, y, min(L, Nz-1)) << bits: layer iz+1 (iz = floor(fz) in
// [-1, Nz-1]) carries BOTH z slices of the trilinear footprint with GL_CLAMP_TO_EDGE in z applied.
// The texel-offset operand (SASS TLD4.R.AOFFI, immediate {1,1}) moves the footprint from {i-1, i} to
// {i, i+1}: the gather coordinate is (float(ix), float(iy)) -- exact, the corner shared by the four
// texels, half a texel from every footprint boundary (immune to the unit's fixed-point rounding).
__device__ __forceinline__ void tld4_pair(cudaTextureObject_t tex, int layer, float x, float y,
                                          uint32_t& t_x0y1, uint32_t& t_x1y1, uint32_t& t_x1y0, uint32_t& t_x0y0)
{
    asm volatile("tld4.r.a2d.v4.u32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}], {1, 1};"
                 : "=r"(t_x0y1), "=r"(t_x1y1), "=r"(t_x1y0), "=r"(t_x0y0)
                 : "l"(tex), "r"(layer), "f"(x), "f"(y));
}
// integer-coordinate texel load, one component used (SASS TLD.LZ)
__device__ __forceinline__ uint32_t tld_layer(cudaTextureObject_t tex, unsigned layer, unsigned x, unsigned y)
{
    uint32_t r, g, b, a;
    asm volatile("tex.a2d.v4.u32.s32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                 : "=r"(r), "=r"(g), "=r"(b), "=r"(a) : "l"(tex), "r"(layer), "r"(x), "r"(y));
    return r;
}

// biased floats 2^23 + v of the two halves of a z-pair texel: one PRMT each (differences of biased
// values are exact, so only the x-low corners are un-biased, inside the lerp's addend).  (A sub-word
// integer->float conversion, I2F.U16 Rx.H0/.H1, would save the un-biasing but runs on the 1/8-rate conversion
// pipe: 8 per sample would make that pipe the bound.)
template <typename T> __device__ __forceinline__ f2 unpack_zpair(uint32_t w);
template <> __device__ __forceinline__ f2 unpack_zpair<uint16_t>(uint32_t w)
{
    return mk2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)));
}
template <> __device__ __forceinline__ f2 unpack_zpair<uint8_t>(uint32_t w)
{
    return mk2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651)));
}

// ---- pieces shared by the two kernels ------------------------------------------------------------

// cartesianToTextureCoord (:175-192) of the sample at (pxy, pz); returns the key of the :118 range test:
// max of the three bit patterns (no -0 / NaN on this path: finite camera, alpha_scale >= 0), > 1.0f's pattern = outside
template <int TCDIV, bool UNIT, int FORM>
__device__ __forceinline__ unsigned tex_coord_key(const FrameConsts& fc, f2 pxy, float pz, f2& txy, float& tz)
{
    const f2 qxy = fadd(pxy, mk2(fc.half_len[0], fc.half_len[1]));
    const float qz = __fadd_rn(pz, fc.half_len[2]);
    float tzq;
    if (UNIT) { txy = qxy; tzq = qz; }                                   // every divisor is exactly 1
    else if (TCDIV == DIV_RECIP_EXACT) { txy = fmul(qxy, mk2(fc.inv_denom[0], fc.inv_denom[1])); tzq = __fmul_rn(qz, fc.inv_denom[2]); }
    else {
        const f2 ixy = mk2(fc.inv_denom[0], fc.inv_denom[1]);
        const f2 q0 = fmul(qxy, ixy);
        const f2 r = ffma(mk2(-fc.denom[0], -fc.denom[1]), q0, qxy);
        txy = ffma(r, ixy, q0);
        tzq = div_by<DIV_MARKSTEIN>(qz, fc.denom[2], fc.inv_denom[2]);
    }
    tz = __fsub_rn(1.0f, tzq);                                           // :185
    if (FORM >= FORM_GENERAL) {
        if (fc.view_top) {                                               // :186-187  (x, 1 - z, y), z already flipped
            const float ty = __fsub_rn(1.0f, tz), qy = hi(txy);
            txy = mk2(lo(txy), ty); tz = qy;
        } else if (fc.view_bottom) {                                     // :188-189  (x, z, 1 - y)
            const float ty = tz, qy = hi(txy);
            txy = mk2(lo(txy), ty); tz = __fsub_rn(1.0f, qy);
        }
    }
    return max(max(__float_as_uint(lo(txy)), __float_as_uint(hi(txy))), __float_as_uint(tz));
}

// window (:122-124) + classification + compositing (:130-132) or MIP (:163-167) of one sample value
template <int WIN, int FORM>
__device__ __forceinline__ void shade_sample(const FrameConsts& fc, const MarchArgs& args, float s, float& C, float& A)
{
    float v;
    if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, fc.frange, fc.inv_frange);
    else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fc.fmin), fc.fmax), fc.fmin), fc.frange, fc.inv_frange);
    if (FORM == FORM_MIP) {
        const float m = __fmul_rn(v, fc.alpha_scale);                    // :164
        if (A < m) { C = m; A = m; }                                     // :165-168
        return;
    }
    float src_a = v;
    bool final_a = false;
    if (FORM == FORM_TF || (FORM >= FORM_GENERAL && fc.use_tf)) {
        // v in [0,1] on this path (ordered window), so the index is in [0,255]; the unsigned min only guards the load
        const unsigned iso = min((unsigned)__float2int_rd(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f)), 255u);
        src_a = __ldg(args.tf_lut + iso);
        final_a = FORM >= FORM_GENERAL && args.lut_final != 0;
    }
    if (FORM >= FORM_GENERAL && fc.is_mip) {
        const float m_rgb = __fmul_rn(v, fc.alpha_scale), m_a = __fmul_rn(src_a, fc.alpha_scale);
        if (A < m_a) { C = m_rgb; A = m_a; }
        return;
    }
    float a = final_a ? src_a : __fmul_rn(src_a, fc.alpha_scale);        // :130
    if (FORM == FORM_GENERAL_OC && !final_a)
        a = (float)(1.0 - pow(1.0 - (double)a, (double)fc.step_scale)); // extension: opacity correction
    const float c = __fmul_rn(v, a);                                     // :131
    const float t = __fsub_rn(1.0f, A);                                  // :132
    const f2 ca_t = fmul(mk2(c, a), splat2(t));
    C = __fadd_rn(C, lo(ca_t));
    A = __fadd_rn(A, hi(ca_t));
}

// ---- empty-space skipping helpers ------------------------------------------------------------------
// is the cell of a sample with base voxel indices (jx-1, jy-1, jz-1) empty?  j in [0, N] (sample inside the box)
__device__ __forceinline__ bool cell_is_empty(const uint32_t* __restrict__ s_mask, const MarchArgs& args,
                                              unsigned jx, unsigned jy, unsigned jz)
{
    const unsigned idx = (jz >> args.cell_shift) * (unsigned)args.cell_nxy + (jy >> args.cell_shift) * (unsigned)args.cell_nx +
                         (jx >> args.cell_shift);
    return (s_mask[idx >> 5] >> (idx & 31u)) & 1u;
}

// Number of steps a ray can advance along one axis and stay inside its cell AND the box, with a 0.05-voxel
// margin.  u = the sample's continuous index coordinate on this axis (the value whose floor picked the voxel),
// j = floor(u) + 1, du = its (approximate) increment per step, [box_lo, box_hi] = the :118 range in the same
// coordinate.  The cell with c = j >> s covers u in [c*2^s - 1, (c+1)*2^s - 1).
__device__ __forceinline__ float axis_steps(float u, unsigned j, float du, int shift, float box_lo, float box_hi)
{
    const unsigned c0 = (j >> shift) << shift;
    const float lo = fmaxf((float)c0 - 1.0f, box_lo) + 0.05f;
    const float hi = fminf((float)(c0 + (1u << shift)) - 1.0f, box_hi) - 0.05f;
    const float dist = du > 0.0f ? hi - u : u - lo;
    const float ad = fabsf(du);
    return ad > 1e-12f ? __fdividef(dist, ad) : 1e9f;
}

// per-ray increments of the three texture-space index coordinates per march step (approximate: they only bound
// leaps); the swizzles follow cartesianToTextureCoord (:175-192)
template <int FORM>
__device__ __forceinline__ void index_increments(const FrameConsts& fc, const float dstep[3], float& dux, float& duy, float& duz)
{
    const float sx = dstep[0] * fc.inv_denom[0], sy = dstep[1] * fc.inv_denom[1], sz = dstep[2] * fc.inv_denom[2];
    float tx = sx, ty = sy, tz = -sz;                                    // (x, y, 1 - z)
    if (FORM >= FORM_GENERAL) {
        if (fc.view_top) { ty = sz; tz = sy; }                           // (x, 1 - (1 - z), y)
        else if (fc.view_bottom) { ty = -sz; tz = -sy; }                 // (x, 1 - z, 1 - y)
    }
    dux = tx * fc.dimf[0]; duy = ty * fc.dimf[1]; duz = tz * fc.dimf[2];
}

// ---- trilinear ------------------------------------------------------------------------------------
struct FetchedPair { uint32_t t01, t11, t10, t00; float wx, wy, wz; };

// :121 (trilinear extension: lerp(a,b,w) = fma(w, b-a, a) in x, then y, then z) + shading of one fetched sample
template <typename T, int WIN, int FORM>
__device__ __forceinline__ void consume_pair(const FrameConsts& fc, const MarchArgs& args, const FetchedPair& f, float& C, float& A)
{
    const f2 B2 = splat2(8388608.0f);
    const f2 wxx = splat2(f.wx), wyy = splat2(f.wy);
    const f2 loA = unpack_zpair<T>(f.t00), hiA = unpack_zpair<T>(f.t10);   // row y   : halves = (z, z+1)
    const f2 loB = unpack_zpair<T>(f.t01), hiB = unpack_zpair<T>(f.t11);   // row y+1
    const f2 cA = ffma(wxx, fsub(hiA, loA), fsub(loA, B2));
    const f2 cB = ffma(wxx, fsub(hiB, loB), fsub(loB, B2));
    const f2 cy = ffma(wyy, fsub(cB, cA), cA);
    const float s = __fmaf_rn(f.wz, __fsub_rn(hi(cy), lo(cy)), lo(cy));
    shade_sample<WIN, FORM>(fc, args, s, C, A);
}

// Software pipeline, DEPTH deep.  The gather coordinates of a sample depend only on `pos`, never on
// fetched data, so the tld4 of sample i+DEPTH-1 is issued BEFORE sample i is unpacked, interpolated
// and composited: DEPTH gathers are in flight per warp when a warp waits (the texture unit returns
// 32 B/clk/SM = 16 cycles per 32-bit tld4 warp instruction; one gather in flight per warp leaves that
// path ~27 % idle).  Ring of DEPTH register sets, loop unrolled DEPTH times: no register moves.
// Scheduling only -- the operation sequence per sample, and hence every bit, is unchanged.  A look-ahead
// sample outside the box is fetched with a clamped layer (x/y clamp is the address mode) and never used.
template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, int FORM, int DEPTH>
__device__ __forceinline__ void march_ray_texpair(const FrameConsts& fc, const MarchArgs& args,
                                                  const float pos0[3], const float dstep[3], float& C, float& A)
{
    static_assert(DEPTH >= 2 && DEPTH <= 4, "pipeline depth");
    const cudaTextureObject_t tex = args.tex;
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const f2 mhalf = splat2(-0.5f);
    const unsigned last_layer = (unsigned)fc.dim[2];

    // texel coordinates, weights, gather.  x / y are only needed as floats (gather coordinate + weight): one
    // FRND.FLOOR each; z is needed as the layer index: F2I.FLOOR + I2FP.  (floorf is exact, so the weights are the
    // same single-rounded differences f - floor(f) as in the oracle.)
    auto fetch = [&](f2 txy, float tz, FetchedPair& f) {
        const f2 fxy = ffma(txy, nxy, mhalf);
        const float fz = __fmaf_rn(tz, nz, -0.5f);
        const float flx = floorf(lo(fxy)), fly = floorf(hi(fxy));
        const int iz = __float2int_rd(fz);
        tld4_pair(tex, (int)min((unsigned)(iz + 1), last_layer), flx, fly, f.t01, f.t11, f.t10, f.t00);
        f.wx = __fsub_rn(lo(fxy), flx); f.wy = __fsub_rn(hi(fxy), fly); f.wz = __fsub_rn(fz, (float)iz);
    };

    f2 txy; float tz;
    if (tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz) > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return;   // :118, first sample
    FetchedPair F[DEPTH];
    unsigned key[DEPTH];                                                 // :118 range key of the sample in each slot
    fetch(txy, tz, F[0]);
    key[0] = 0u;
#pragma unroll
    for (int k = 1; k < DEPTH - 1; ++k) {                               // prologue: samples 1 .. DEPTH-2
        pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz);                    // :136
        key[k] = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);
        fetch(txy, tz, F[k]);
    }
    int iter = 0;
    // one pass of the unrolled ring; false = the ray is finished
    auto pass = [&]() -> bool {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) {
            const int fill = (k + DEPTH - 1) % DEPTH, next = (k + 1) % DEPTH;
            // look ahead: sample (iter + k) + DEPTH-1
            pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz);                // :136
            key[fill] = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);
            fetch(txy, tz, F[fill]);
            // finish sample iter + k
            consume_pair<T, WIN, FORM>(fc, args, F[k], C, A);
            // :118 of sample iter + k + 1 (the `dest.a > 0.99` break of :134 is subsumed by it: only `pos` changes in between)
            if (key[next] > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return false;
            if (!NOCAP && iter + k + 1 >= 10000) return false;           // :115
        }
        iter += DEPTH;
        return true;
    };
    while (pass()) {}
}

// The same march with empty-space skipping (see the file header).  The pipelined loop is kept as it is; the cell
// map is consulted only at CHECKPOINTS: the first sample and then every (skip_check_mask+1)-th pass of the unrolled
// loop -- a WARP-UNIFORM condition, so lanes never wait for each other's checkpoints.  At a checkpoint `examine`
// looks at the front sample's cell.  Non-empty: carry on.  Empty: leap -- the number of further samples certain to
// stay in the cell (a LOWER bound, 0.05 voxel inside the cell and the box) is skipped by performing exactly that
// many position updates (:136), then the landing sample is examined in full (box test :118, cell test) and so on
// until a non-empty cell or the end of the box.  Skipped samples never enter the ring; they contribute exactly 0,
// so consuming the ring in order reproduces the reference's accumulation sequence.  (Samples of an empty cell
// met between two checkpoints are simply processed: always exact, only slower.)

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, int FORM, int DEPTH>
__device__ __forceinline__ void march_ray_texpair_skip(const FrameConsts& fc, const MarchArgs& args, const uint32_t* __restrict__ s_mask,
                                                       const float pos0[3], const float dstep[3], float& C, float& A)
{
    static_assert(DEPTH >= 2 && DEPTH <= 4, "pipeline depth");
    const cudaTextureObject_t tex = args.tex;
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const f2 mhalf = splat2(-0.5f);
    const unsigned last_layer = (unsigned)fc.dim[2];
    int jf = 0;                                                          // index of the front sample (:115; only read when !NOCAP)

    // checkpoint: leap over the empty cells in front of the ray; on return (txy, tz, key) describe the new front sample
    auto examine = [&](f2& txy, float& tz, unsigned& key) {
        while (key <= 0x3F800000u) {
            const f2 fxy = ffma(txy, nxy, mhalf);
            const float fz = __fmaf_rn(tz, nz, -0.5f);
            const int ix = __float2int_rd(lo(fxy)), iy = __float2int_rd(hi(fxy)), iz = __float2int_rd(fz);
            if (!cell_is_empty(s_mask, args, (unsigned)(ix + 1), (unsigned)(iy + 1), (unsigned)(iz + 1))) return;
            // empty cell: this sample and the next k are certain to contribute nothing
            float dux, duy, duz;
            index_increments<FORM>(fc, dstep, dux, duy, duz);
            const float st = fminf(fminf(axis_steps(lo(fxy), (unsigned)(ix + 1), dux, args.cell_shift, -0.5f, fc.dimf[0] - 0.5f),
                                         axis_steps(hi(fxy), (unsigned)(iy + 1), duy, args.cell_shift, -0.5f, fc.dimf[1] - 0.5f)),
                                   axis_steps(fz, (unsigned)(iz + 1), duz, args.cell_shift, -0.5f, fc.dimf[2] - 0.5f));
            const int k = st > 0.0f ? (int)fminf(st, 65535.0f) : 0;
            for (int n = 0; n <= k; ++n) { pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz); }                    // :136, k+1 times
            if (!NOCAP) jf += k + 1;
            key = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);                                     // :118 of the landing sample
        }
    };
    // the front stage: box test of the sample at `pos`, (checkpoint,) gather -- unconditional, like the plain loop;
    // returns the :118 range key of the sample now in the slot (> 1.0f's bit pattern: no such sample)
    auto front = [&](FetchedPair& f, bool checkpoint) -> unsigned {
        f2 txy; float tz;
        unsigned key = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);                                // :118
        if (checkpoint) examine(txy, tz, key);
        if (!NOCAP && jf >= 10000) key = 0xFFFFFFFFu;                                                         // :115
        const f2 fxy = ffma(txy, nxy, mhalf);
        const float fz = __fmaf_rn(tz, nz, -0.5f);
        const float flx = floorf(lo(fxy)), fly = floorf(hi(fxy));
        const int iz = __float2int_rd(fz);
        tld4_pair(tex, (int)min((unsigned)(iz + 1), last_layer), flx, fly, f.t01, f.t11, f.t10, f.t00);
        f.wx = __fsub_rn(lo(fxy), flx); f.wy = __fsub_rn(hi(fxy), fly); f.wz = __fsub_rn(fz, (float)iz);
        return key;
    };
    auto advance = [&]() { pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz); if (!NOCAP) ++jf; };                // :136

    if (__float_as_uint(A) >= 0x3F733333u) return;
    FetchedPair F[DEPTH];
    unsigned key[DEPTH];
    key[0] = front(F[0], true);
    if (key[0] > 0x3F800000u) return;
#pragma unroll
    for (int k = 1; k < DEPTH - 1; ++k) { advance(); key[k] = front(F[k], false); }
    int it = 1;
    auto pass = [&]() -> bool {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) {
            const int fill = (k + DEPTH - 1) % DEPTH, next = (k + 1) % DEPTH;
            advance();
            key[fill] = front(F[fill], k == 0 && (it & args.skip_check_mask) == 0);
            consume_pair<T, WIN, FORM>(fc, args, F[k], C, A);
            if (key[next] > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return false;
        }
        ++it;
        return true;
    };
    while (pass()) {}
}

// A warp covers an 8 x 4 pixel patch; a CTA of CTAW warps covers (8 * min(CTAW,4)) x (4 * CTAW / min(CTAW,4)) pixels:
// 8 warps = 32 x 8, 4 warps = 32 x 4, 2 warps = 16 x 4.  MINW = resident WARPS per SM the register budget is sized for.  Smaller CTAs return their registers sooner (a CTA lives as long as its slowest
// warp) and balance the tail of small grids (multi-GPU partitions) at a finer grain.
template <int CTAW, int WX_ = (CTAW < 4 ? CTAW : 4)> struct CtaShape {
    static constexpr int WX = WX_, WY = CTAW / WX, PX = 8 * WX, PY = 4 * WY, THREADS = 32 * CTAW;
};

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, int FORM, int DEPTH, bool SKIP, int MINW, int CTAW = 8, int CTAWX = (CTAW < 4 ? CTAW : 4)>
__global__ void __launch_bounds__(32 * CTAW, MINW / CTAW)
march_texpair_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ MarchArgs args)
{
    typedef CtaShape<CTAW, CTAWX> S;
    extern __shared__ uint32_t s_mask[];
    if (SKIP) {                                                          // stage the empty-cell bit map (a few KB)
        for (int i = threadIdx.x; i < args.cell_words; i += S::THREADS) s_mask[i] = __ldg(args.cell_bits + i);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned bx = blockIdx.x, by = blockIdx.y;
    if (args.cta_order) { const unsigned t = __ldg(args.cta_order + blockIdx.y * gridDim.x + blockIdx.x); bx = t & 0xffffu; by = t >> 16; }
    const int px = bx * S::PX + (warp % S::WX) * 8 + (lane & 7);
    const int lrow = fc.row0 + by * S::PY + (warp / S::WX) * 4 + (lane >> 3);
    const int py = owned_row_to_global(fc, lrow);
    if (px < fc.W && lrow < args.local_rows && py < fc.H) {
        const RaySetup r = setup_ray(fc, px, py);
        float C = 0.0f, A = 0.0f;
        if (r.hit) {
            float pos[3], ds[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));   // :107, :114
                ds[i] = __fmul_rn(r.dir[i], fc.step);                                                                    // :136
            }
            if (SKIP) march_ray_texpair_skip<T, TCDIV, WIN, UNIT, NOCAP, FORM, DEPTH>(fc, args, s_mask, pos, ds, C, A);
            else march_ray_texpair<T, TCDIV, WIN, UNIT, NOCAP, FORM, DEPTH>(fc, args, pos, ds, C, A);
        }
        const int orow = fc.compact ? lrow : py;
        reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
    }
    if (args.peer_arrive) {                                              // fused hand-off (uniform branch)
        __syncthreads();                                                 // this CTA's pixel stores are ordered before ...
        if (threadIdx.x == 0) {
            __threadfence_system();                                      // ... this (cumulative) system-scope fence
            if (atomicAdd(args.done_counter, 1u) == args.grid_ctas - 1u) {
                *args.done_counter = 0u;                                 // re-arm for the next frame (stream order)
                __threadfence_system();
                atomicAdd_system(args.peer_arrive, 1u);
            }
        }
    }
}

// ---- nearest --------------------------------------------------------------------------------------
// i = clamp(floor(u * N), 0, N-1) per axis (VolumeRenderer.cs:121 + GL_CLAMP_TO_EDGE): the clamp is an
// unsigned min, which also makes the look-ahead fetch safe when that sample lies outside the box.
template <int TCDIV, int WIN, bool UNIT, bool NOCAP, int FORM, int DEPTH>
__device__ __forceinline__ void march_ray_nearest(const FrameConsts& fc, const MarchArgs& args,
                                                  const float pos0[3], const float dstep[3], float& C, float& A)
{
    static_assert(DEPTH >= 2 && DEPTH <= 4, "pipeline depth");
    const cudaTextureObject_t tex = args.tex;
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const unsigned mx = (unsigned)fc.dim[0] - 1u, my = (unsigned)fc.dim[1] - 1u, mz = (unsigned)fc.dim[2] - 1u;

    auto fetch = [&](f2 txy, float tz) -> uint32_t {
        const f2 uxy = fmul(txy, nxy);
        const float uz = __fmul_rn(tz, nz);
        const unsigned ix = min((unsigned)__float2int_rd(lo(uxy)), mx);
        const unsigned iy = min((unsigned)__float2int_rd(hi(uxy)), my);
        const unsigned iz = min((unsigned)__float2int_rd(uz), mz);
        return tld_layer(tex, iz, ix, iy);
    };
    auto consume = [&](uint32_t texel) {                                 // :121-132
        const float s = __fsub_rn(__uint_as_float(0x4B000000u | texel), 8388608.0f);
        shade_sample<WIN, FORM>(fc, args, s, C, A);
    };

    f2 txy; float tz;
    if (tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz) > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return;
    uint32_t F[DEPTH];
    unsigned key[DEPTH];
    F[0] = fetch(txy, tz);
    key[0] = 0u;
#pragma unroll
    for (int k = 1; k < DEPTH - 1; ++k) {
        pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz);
        key[k] = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);
        F[k] = fetch(txy, tz);
    }
    int iter = 0;
    auto pass = [&]() -> bool {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) {
            const int fill = (k + DEPTH - 1) % DEPTH, next = (k + 1) % DEPTH;
            pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz);                // :136
            key[fill] = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);
            F[fill] = fetch(txy, tz);
            consume(F[k]);
            if (key[next] > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return false;
            if (!NOCAP && iter + k + 1 >= 10000) return false;
        }
        iter += DEPTH;
        return true;
    };
    while (pass()) {}
}

// nearest filter with empty-space skipping: same checkpoint / leap logic as march_ray_texpair_skip; the index
// coordinate is u = t*N (the voxel is floor(u), clamped), the :118 range is u in [0, N]
template <int TCDIV, int WIN, bool UNIT, bool NOCAP, int FORM, int DEPTH>
__device__ __forceinline__ void march_ray_nearest_skip(const FrameConsts& fc, const MarchArgs& args, const uint32_t* __restrict__ s_mask,
                                                       const float pos0[3], const float dstep[3], float& C, float& A)
{
    static_assert(DEPTH >= 2 && DEPTH <= 4, "pipeline depth");
    const cudaTextureObject_t tex = args.tex;
    f2 pxy = mk2(pos0[0], pos0[1]);
    float pz = pos0[2];
    const f2 dxy = mk2(dstep[0], dstep[1]);
    const float dz = dstep[2];
    const f2 nxy = mk2(fc.dimf[0], fc.dimf[1]);
    const float nz = fc.dimf[2];
    const unsigned mx = (unsigned)fc.dim[0] - 1u, my = (unsigned)fc.dim[1] - 1u, mz = (unsigned)fc.dim[2] - 1u;
    int jf = 0;

    auto examine = [&](f2& txy, float& tz, unsigned& key) {
        while (key <= 0x3F800000u) {
            const f2 uxy = fmul(txy, nxy);
            const float uz = __fmul_rn(tz, nz);
            const unsigned ix = min((unsigned)__float2int_rd(lo(uxy)), mx);
            const unsigned iy = min((unsigned)__float2int_rd(hi(uxy)), my);
            const unsigned iz = min((unsigned)__float2int_rd(uz), mz);
            if (!cell_is_empty(s_mask, args, ix + 1u, iy + 1u, iz + 1u)) return;
            float dux, duy, duz;
            index_increments<FORM>(fc, dstep, dux, duy, duz);
            const float st = fminf(fminf(axis_steps(lo(uxy), ix + 1u, dux, args.cell_shift, 0.0f, fc.dimf[0]),
                                         axis_steps(hi(uxy), iy + 1u, duy, args.cell_shift, 0.0f, fc.dimf[1])),
                                   axis_steps(uz, iz + 1u, duz, args.cell_shift, 0.0f, fc.dimf[2]));
            const int k = st > 0.0f ? (int)fminf(st, 65535.0f) : 0;
            for (int n = 0; n <= k; ++n) { pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz); }                    // :136, k+1 times
            if (!NOCAP) jf += k + 1;
            key = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);
        }
    };
    auto front = [&](uint32_t& texel, bool checkpoint) -> unsigned {
        f2 txy; float tz;
        unsigned key = tex_coord_key<TCDIV, UNIT, FORM>(fc, pxy, pz, txy, tz);                                // :118
        if (checkpoint) examine(txy, tz, key);
        if (!NOCAP && jf >= 10000) key = 0xFFFFFFFFu;                                                         // :115
        const f2 uxy = fmul(txy, nxy);
        const float uz = __fmul_rn(tz, nz);
        const unsigned ix = min((unsigned)__float2int_rd(lo(uxy)), mx);
        const unsigned iy = min((unsigned)__float2int_rd(hi(uxy)), my);
        const unsigned iz = min((unsigned)__float2int_rd(uz), mz);
        texel = tld_layer(tex, iz, ix, iy);
        return key;
    };
    auto advance = [&]() { pxy = fadd(pxy, dxy); pz = __fadd_rn(pz, dz); if (!NOCAP) ++jf; };
    auto consume = [&](uint32_t texel) {
        const float s = __fsub_rn(__uint_as_float(0x4B000000u | texel), 8388608.0f);
        shade_sample<WIN, FORM>(fc, args, s, C, A);
    };

    if (__float_as_uint(A) >= 0x3F733333u) return;
    uint32_t F[DEPTH];
    unsigned key[DEPTH];
    key[0] = front(F[0], true);
    if (key[0] > 0x3F800000u) return;
#pragma unroll
    for (int k = 1; k < DEPTH - 1; ++k) { advance(); key[k] = front(F[k], false); }
    int it = 1;
    auto pass = [&]() -> bool {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) {
            const int fill = (k + DEPTH - 1) % DEPTH, next = (k + 1) % DEPTH;
            advance();
            key[fill] = front(F[fill], k == 0 && (it & args.skip_check_mask) == 0);
            consume(F[k]);
            if (key[next] > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) return false;
        }
        ++it;
        return true;
    };
    while (pass()) {}
}

template <int TCDIV, int WIN, bool UNIT, bool NOCAP, int FORM, int DEPTH, bool SKIP, int MINW, int CTAW = 8, int CTAWX = (CTAW < 4 ? CTAW : 4)>
__global__ void __launch_bounds__(32 * CTAW, MINW / CTAW)
march_nearest_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ MarchArgs args)
{
    typedef CtaShape<CTAW, CTAWX> S;
    extern __shared__ uint32_t s_mask[];
    if (SKIP) {
        for (int i = threadIdx.x; i < args.cell_words; i += S::THREADS) s_mask[i] = __ldg(args.cell_bits + i);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned bx = blockIdx.x, by = blockIdx.y;
    if (args.cta_order) { const unsigned t = __ldg(args.cta_order + blockIdx.y * gridDim.x + blockIdx.x); bx = t & 0xffffu; by = t >> 16; }
    const int px = bx * S::PX + (warp % S::WX) * 8 + (lane & 7);
    const int lrow = fc.row0 + by * S::PY + (warp / S::WX) * 4 + (lane >> 3);
    const int py = owned_row_to_global(fc, lrow);
    if (px < fc.W && lrow < args.local_rows && py < fc.H) {
        const RaySetup r = setup_ray(fc, px, py);
        float C = 0.0f, A = 0.0f;
        if (r.hit) {
            float pos[3], ds[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));
                ds[i] = __fmul_rn(r.dir[i], fc.step);
            }
            if (SKIP) march_ray_nearest_skip<TCDIV, WIN, UNIT, NOCAP, FORM, DEPTH>(fc, args, s_mask, pos, ds, C, A);
            else march_ray_nearest<TCDIV, WIN, UNIT, NOCAP, FORM, DEPTH>(fc, args, pos, ds, C, A);
        }
        const int orow = fc.compact ? lrow : py;
        reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);
    }
    if (args.peer_arrive) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(args.done_counter, 1u) == args.grid_ctas - 1u) {
                *args.done_counter = 0u;
                __threadfence_system();
                atomicAdd_system(args.peer_arrive, 1u);
            }
        }
    }
}

}  // namespace vr
