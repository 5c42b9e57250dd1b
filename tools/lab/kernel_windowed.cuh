// kernel_windowed.cuh -- persistent screen-tile march through TMA-staged shared-memory windows.
//
// One CTA = one producer warp + 256 consumer threads (a 16x16-pixel screen tile, one ray
// each).  CTAs are persistent: they pull screen tiles from an atomic counter until the frame
// is done.  For a tile whose rays travel mainly along the volume's z axis the producer
// thread walks a sequence of WINDOWS -- BZ consecutive z-slices of the edge-replicated
// volume, each slice a BX x BY box whose origin follows the tile's frustum (integer shear
// per slice) -- and stages window w+1 into shared memory with BZ cp.async.bulk.tensor (TMA)
// tile loads on an mbarrier while the consumers march window w out of the other buffer.
// A consumer takes every sample whose 2x2x2 neighbourhood lies inside the current window
// from shared memory (8 LDS.U16), parks when its next sample is beyond the window in travel
// direction, and falls back to a global-memory fetch for any other sample (window too
// small for the tile, ray grazing the window side) -- so the image never depends on the
// window geometry, only the speed does.  The arithmetic is the same correctly-rounded
// sequence as everywhere else (march_device.cuh); results are bit-identical to the oracle.
//
// Tiles that are not z-major (or frames the fast path does not cover) run the plain global
// march inside the same persistent kernel.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "kernel_lab.cuh"

namespace vr {

constexpr int WT_TILE = 16;                 // screen tile edge, pixels
constexpr int WT_CONSUMERS = WT_TILE * WT_TILE;
constexpr int WT_THREADS = WT_CONSUMERS + 32;
constexpr int WT_STAGES = 2;

template <typename T> struct WinGeom;
// TMA tile loads need the innermost start coordinate 16-byte aligned (measured:
// profiles/microbench/tma_probe.cu), so a window cannot shear in x: its x origin is aligned
// down to XALIGN voxels and BX absorbs alignment slack + the x drift over BZ slices.  The y
// origin is free and follows the frustum with an integer shear per slice.
template <> struct WinGeom<uint16_t> { static constexpr int BX = 40, BY = 24, BZ = 10, XALIGN = 8; };
template <> struct WinGeom<uint8_t>  { static constexpr int BX = 48, BY = 24, BZ = 10, XALIGN = 16; };

struct WindowDesc {          // written by the producer, read by the consumers
    int a;                   // padded z index of slice 0
    int ox0, oy0;            // lateral origin of slice 0 (padded indices)
    int shx, shy;            // origin of slice k = (ox0 + k*shx, oy0 + k*shy)
};

struct TileInfo {
    int tile;                // tile id, or -1 when the frame is finished
    int any_hit;
    int zs_min, zs_max;      // extreme first-sample base z index among the tile's rays
    int zmajor, sgn;         // traversal class of the tile (from its centre ray), travel direction in z
};

struct WindowedArgs {
    const void* vol;
    uint32_t pitch, slice_lo;
    float* out;
    int local_rows;
    int tiles_x, tiles_y;
    unsigned int* tile_counter;
    unsigned long long* stats;   // optional: [0] smem samples, [1] fallback samples, [2] windows
};

// ---------------------------------------------------------------- PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// affine map from a world position to padded float texel coordinates (f + 1):
// F = ((p + half)/D [z: 1 - .]) * N - 0.5 + 1.  Geometry only (window placement); the march
// itself never uses it.
struct IndexMap {
    float ax, bx, ay, by, az, bz;
    __device__ __forceinline__ void init(const FrameConsts& fc)
    {
        ax = fc.dimf[0] / fc.denom[0]; bx = fc.half_len[0] * ax + 0.5f;
        ay = fc.dimf[1] / fc.denom[1]; by = fc.half_len[1] * ay + 0.5f;
        az = -fc.dimf[2] / fc.denom[2]; bz = fc.dimf[2] - fc.half_len[2] * fc.dimf[2] / fc.denom[2] + 0.5f;
    }
};

struct CornerRays {          // producer-private: the tile's four corner pixels + the centre, in index space
    float o[3];              // F(origin)
    float g[5][3];           // dF/dt for corner rays 0..3 and the centre ray 4
};

// producer: geometry of the window whose slice 0 is padded z index `a`.  Conservative
// estimate only -- consumers test every sample against the window and fall back to global
// memory, so a poor plan costs speed, never correctness.
template <typename T>
__device__ __noinline__ void plan_window(const CornerRays& cr, int a, WindowDesc& wd)
{
    constexpr int BZ = WinGeom<T>::BZ, BX = WinGeom<T>::BX, BY = WinGeom<T>::BY;
    float inv_gz[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) inv_gz[i] = 1.0f / cr.g[i][2];
    // integer shear per slice from the centre ray
    const float sx = cr.g[4][0] * inv_gz[4], sy = cr.g[4][1] * inv_gz[4];      // d(fx)/d(fz), d(fy)/d(fz)
    (void)sx;
    const int shx = 0, shy = max(-1, min(1, __float2int_rn(sy)));
    // plane j (fz = a + j) bounds the samples that touch slices j-1, j, j+1; slide a 3-plane window
    float pxlo[3], pxhi[3], pylo[3], pyhi[3];
    int ox0 = 0x7fffffff, oy0 = 0x7fffffff, ex = -0x7fffffff, ey = -0x7fffffff;
    for (int j = 0; j <= BZ; ++j) {
        const int jj = min(j, BZ - 1);
        const float c = (float)(a + jj);
        float xl = 3.0e38f, xh = -3.0e38f, yl = 3.0e38f, yh = -3.0e38f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float t = (c - cr.o[2]) * inv_gz[i];
            const float x = cr.o[0] + t * cr.g[i][0], y = cr.o[1] + t * cr.g[i][1];
            xl = fminf(xl, x); xh = fmaxf(xh, x); yl = fminf(yl, y); yh = fmaxf(yh, y);
        }
        pxlo[j % 3] = xl; pxhi[j % 3] = xh; pylo[j % 3] = yl; pyhi[j % 3] = yh;
        if (j >= 1) {
            const int k = j - 1;                                 // slice k: planes k-1, k, k+1
            float l = fminf(pxlo[k % 3], pxlo[j % 3]), h = fmaxf(pxhi[k % 3], pxhi[j % 3]);
            float m = fminf(pylo[k % 3], pylo[j % 3]), n = fmaxf(pyhi[k % 3], pyhi[j % 3]);
            if (k >= 1) { l = fminf(l, pxlo[(k - 1) % 3]); h = fmaxf(h, pxhi[(k - 1) % 3]); m = fminf(m, pylo[(k - 1) % 3]); n = fmaxf(n, pyhi[(k - 1) % 3]); }
            // one texel of slack each side for fp32 drift, +1 for the upper neighbour
            ox0 = min(ox0, __float2int_rd(l) - 1 - k * shx); ex = max(ex, __float2int_rd(h) + 2 - k * shx);
            oy0 = min(oy0, __float2int_rd(m) - 1 - k * shy); ey = max(ey, __float2int_rd(n) + 2 - k * shy);
        }
    }
    // centre the slack when the box is wider than the footprint
    const int wx = ex - ox0 + 1, wy = ey - oy0 + 1;
    if (wx < BX) ox0 -= (BX - wx) / 2;
    if (wy < BY) oy0 -= (BY - wy) / 2;
    {   // align the x origin down (floor, also for negative coordinates); keep the footprint inside when possible
        constexpr int XA = WinGeom<T>::XALIGN;
        int al = ox0 & ~(XA - 1);
        if (ex - al + 1 > BX && wx <= BX) al = (ex - BX + 1 + XA - 1) & ~(XA - 1);      // slide right if that still covers the low edge
        ox0 = al;
    }
    wd.a = a; wd.ox0 = ox0; wd.oy0 = oy0; wd.shx = shx; wd.shy = shy;
}

// the trilinear sample from 8 biased texels (2^23 + v); identical arithmetic for both sources
__device__ __forceinline__ float lerp8(const float b[8], float wx, float wy, float wz)
{
    const float B = 8388608.0f;
    const float c00 = __fmaf_rn(wx, __fsub_rn(b[1], b[0]), __fsub_rn(b[0], B));
    const float c10 = __fmaf_rn(wx, __fsub_rn(b[3], b[2]), __fsub_rn(b[2], B));
    const float c01 = __fmaf_rn(wx, __fsub_rn(b[5], b[4]), __fsub_rn(b[4], B));
    const float c11 = __fmaf_rn(wx, __fsub_rn(b[7], b[6]), __fsub_rn(b[6], B));
    const float c0 = __fmaf_rn(wy, __fsub_rn(c10, c00), c00);
    const float c1 = __fmaf_rn(wy, __fsub_rn(c11, c01), c01);
    return __fmaf_rn(wz, __fsub_rn(c1, c0), c0);
}

// ld.shared.{u8,u16} into a 32-bit register (zero extended), byte offset folded into the address
template <typename T> __device__ __forceinline__ uint32_t lds_texel(uint32_t addr, int byte_off);
template <> __device__ __forceinline__ uint32_t lds_texel<uint16_t>(uint32_t addr, int byte_off)
{
    uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr + byte_off)); return v;
}
template <> __device__ __forceinline__ uint32_t lds_texel<uint8_t>(uint32_t addr, int byte_off)
{
    uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr + byte_off)); return v;
}

// rare path: the 2x2x2 neighbourhood straight from the padded volume in global memory
template <typename T>
__device__ __noinline__ float sample_global_fallback(const T* __restrict__ vol, uint32_t pitch, uint32_t slice,
                                                      int jx, int jy, int jz, float wx, float wy, float wz)
{
    const uint32_t e00 = (uint32_t)jz * slice + ((uint32_t)jy * pitch + (uint32_t)jx);
    const uint32_t e10 = e00 + pitch, e01 = e00 + slice, e11 = e01 + pitch;
    float b[8];
    b[0] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e00)); b[1] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e00 + 1));
    b[2] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e10)); b[3] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e10 + 1));
    b[4] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e01)); b[5] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e01 + 1));
    b[6] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e11)); b[7] = __uint_as_float(0x4B000000u | (uint32_t)__ldg(vol + e11 + 1));
    return lerp8(b, wx, wy, wz);
}

// tiles that are not z-major: the plain global march, kept out of line
template <typename T, int TCDIV, int WIN>
__device__ __noinline__ void march_global_tile(const FrameConsts& fc, const T* __restrict__ vol, uint32_t pitch, uint32_t slice,
                                               const float pos[3], const float dstep[3], float& C, float& A)
{
    f1 p[3] = {{pos[0]}, {pos[1]}, {pos[2]}}, d[3] = {{dstep[0]}, {dstep[1]}, {dstep[2]}};
    f1 c{0.0f}, a{0.0f};
    int iter = 0;
    march_lanes<T, VR_FILTER_TRILINEAR, TCDIV, WIN, FLOOR_XU1, f1>(fc, vol, pitch, slice, p, d, c, a, iter);
    C = c.v; A = a.v;
}

__device__ __noinline__ void setup_ray_ool(const FrameConsts& fc, int px, int py, float pos[3], float dstep[3], bool& hit)
{
    const RaySetup r = setup_ray(fc, px, py);
    hit = r.hit;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        pos[i] = __fadd_rn(__fadd_rn(r.org[i], __fmul_rn(r.dir[i], r.t_min)), __fmul_rn(r.dir[i], 0.000001f));   // :107,:114
        dstep[i] = __fmul_rn(r.dir[i], fc.step);                                                                  // :136
    }
}

// issuer: corner / centre rays of the tile in index space + traversal class
__device__ __noinline__ void tile_geometry(const FrameConsts& fc, const IndexMap& im, int tile_x, int tile_y, int local_rows,
                                           CornerRays& cr, int& zmajor, int& sgn)
{
    const int x0 = tile_x * WT_TILE, x1 = min(x0 + WT_TILE - 1, fc.W - 1);
    const int r0 = tile_y * WT_TILE, r1 = min(r0 + WT_TILE - 1, local_rows - 1);
    const int y0 = min(owned_row_to_global(fc, r0), fc.H - 1), y1 = min(owned_row_to_global(fc, r1), fc.H - 1);
    const int cxs[5] = {x0, x1, x0, x1, (x0 + x1) >> 1}, cys[5] = {y0, y0, y1, y1, (y0 + y1) >> 1};
    for (int i = 0; i < 5; ++i) {
        const RaySetup r = setup_ray(fc, cxs[i], cys[i]);
        if (i == 0) { cr.o[0] = im.ax * r.org[0] + im.bx; cr.o[1] = im.ay * r.org[1] + im.by; cr.o[2] = im.az * r.org[2] + im.bz; }
        cr.g[i][0] = im.ax * r.dir[0]; cr.g[i][1] = im.ay * r.dir[1]; cr.g[i][2] = im.az * r.dir[2];
    }
    const float gx = fabsf(cr.g[4][0]), gy = fabsf(cr.g[4][1]), gz = cr.g[4][2];
    zmajor = (fabsf(gz) >= 1.05f * fmaxf(gx, gy) && (y1 - y0) < 2 * WT_TILE) ? 1 : 0;
    sgn = gz > 0.0f ? 1 : -1;
}

// keep a loop invariant in a register: the compiler may not re-derive it inside the loop
__device__ __forceinline__ void pin(int& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(float& v) { asm volatile("" : "+f"(v)); }

struct RayState {
    float pos[3], dstep[3];
    float C, A;
    int iter;
    int active;              // 1 while the ray still has samples to take
    unsigned n_smem, n_fallback;
};

// One ray through one window.  Out of line on purpose: register allocation of this hot loop
// is then independent of the (register-hungry, rarely executed) tile set-up code.
template <typename T, int TCDIV, int WIN, bool STATS>
__device__ __noinline__ void march_window(const FrameConsts& fc, const T* __restrict__ vol, uint32_t pitch, uint32_t slice,
                                          uint32_t win_base, int wa, int wox, int woy, int wshy, int sgn, RayState& rs)
{
    constexpr int BX = WinGeom<T>::BX, BY = WinGeom<T>::BY, BZ = WinGeom<T>::BZ;
    constexpr int ES = (int)sizeof(T);
    const int dz1 = (BX * BY - wshy * BX) * ES;                // byte offset of the z+1 texel (y-sheared slice)
    wa -= 1; wox -= 1; woy -= 1;                               // fold the "+1" of the padded index
    // rows ly (slice lz) and ly - shy (slice lz+1) must both lie in [0, BY-2]
    woy += wshy > 0 ? wshy : 0;
    const unsigned yspan = (unsigned)(BY - 2 - (wshy < 0 ? -wshy : wshy));
    const int yfix = wshy > 0 ? wshy : 0;                      // ly below is relative to the shifted origin
    float px = rs.pos[0], py = rs.pos[1], pz = rs.pos[2];
    const float dx = rs.dstep[0], dy = rs.dstep[1], dzs = rs.dstep[2];
    float C = rs.C, A = rs.A;
    int iter = rs.iter;
    int active = 1;
    unsigned n_smem = 0, n_fallback = 0;
    const float h0 = fc.half_len[0], h1 = fc.half_len[1], h2 = fc.half_len[2];
    const float n0 = fc.dimf[0], n1 = fc.dimf[1], n2 = fc.dimf[2];
    const float alpha = fc.alpha_scale, frange = fc.frange, inv_frange = fc.inv_frange, fmin_ = fc.fmin, fmax_ = fc.fmax;
    const float d0 = fc.denom[0], d1 = fc.denom[1], d2 = fc.denom[2];
    const float i0 = fc.inv_denom[0], i1 = fc.inv_denom[1], i2 = fc.inv_denom[2];
    for (; iter < 10000; ++iter) {
        // cartesianToTextureCoord :175-192
        const float tx = div_by<TCDIV>(__fadd_rn(px, h0), d0, i0);
        const float ty = div_by<TCDIV>(__fadd_rn(py, h1), d1, i1);
        const float tz = __fsub_rn(1.0f, div_by<TCDIV>(__fadd_rn(pz, h2), d2, i2));
        const unsigned m = max(max(__float_as_uint(tx), __float_as_uint(ty)), __float_as_uint(tz));
        if (m > 0x3F800000u || __float_as_uint(A) >= 0x3F733333u) { active = 0; break; }     // :118
        const float fx = __fmaf_rn(tx, n0, -0.5f), fy = __fmaf_rn(ty, n1, -0.5f), fz = __fmaf_rn(tz, n2, -0.5f);
        const int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
        const float wx = __fsub_rn(fx, (float)ix), wy = __fsub_rn(fy, (float)iy), wz = __fsub_rn(fz, (float)iz);
        const int lz = iz - wa;
        const int lx = ix - wox, ly = iy - woy - lz * wshy;
        float s;
        if ((unsigned)lz <= (unsigned)(BZ - 2) && (unsigned)lx <= (unsigned)(BX - 2) && (unsigned)ly <= yspan) {
            const uint32_t p0 = win_base + (uint32_t)(((lz * BY + ly + yfix) * BX + lx) * ES);
            const uint32_t p1 = p0 + (uint32_t)dz1;
            float b[8];
            b[0] = __uint_as_float(0x4B000000u | lds_texel<T>(p0, 0));        b[1] = __uint_as_float(0x4B000000u | lds_texel<T>(p0, ES));
            b[2] = __uint_as_float(0x4B000000u | lds_texel<T>(p0, BX * ES));  b[3] = __uint_as_float(0x4B000000u | lds_texel<T>(p0, BX * ES + ES));
            b[4] = __uint_as_float(0x4B000000u | lds_texel<T>(p1, 0));        b[5] = __uint_as_float(0x4B000000u | lds_texel<T>(p1, ES));
            b[6] = __uint_as_float(0x4B000000u | lds_texel<T>(p1, BX * ES));  b[7] = __uint_as_float(0x4B000000u | lds_texel<T>(p1, BX * ES + ES));
            s = lerp8(b, wx, wy, wz);
            if (STATS) ++n_smem;
        } else {
            if (sgn > 0 ? lz > BZ - 2 : lz < 0) break;           // beyond this window in travel direction: park
            s = sample_global_fallback<T>(vol, pitch, slice, ix + 1, iy + 1, iz + 1, wx, wy, wz);
            if (STATS) ++n_fallback;
        }
        float v;                                                // :122-124
        if (WIN == WIN_COVERS0) v = div_by<DIV_MARKSTEIN>(s, frange, inv_frange);
        else v = div_by<DIV_MARKSTEIN>(__fsub_rn(fminf(fmaxf(s, fmin_), fmax_), fmin_), frange, inv_frange);
        const float a = __fmul_rn(v, alpha);                    // :130-132
        const float c = __fmul_rn(v, a);
        const float t = __fsub_rn(1.0f, A);
        C = __fadd_rn(C, __fmul_rn(c, t));
        A = __fadd_rn(A, __fmul_rn(a, t));
        px = __fadd_rn(px, dx);                                 // :136
        py = __fadd_rn(py, dy);
        pz = __fadd_rn(pz, dzs);
    }
    if (iter >= 10000) active = 0;
    rs.pos[0] = px; rs.pos[1] = py; rs.pos[2] = pz;
    rs.C = C; rs.A = A; rs.iter = iter; rs.active = active;
    if (STATS) { rs.n_smem += n_smem; rs.n_fallback += n_fallback; }
}

template <typename T, int TCDIV, int WIN, bool STATS>
__global__ void __launch_bounds__(WT_THREADS)
march_windowed_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ WindowedArgs args,
                      const __grid_constant__ CUtensorMap tmap)
{
    constexpr int BX = WinGeom<T>::BX, BY = WinGeom<T>::BY, BZ = WinGeom<T>::BZ;
    constexpr int SLICE_ELEMS = BX * BY, WIN_ELEMS = SLICE_ELEMS * BZ;
    constexpr unsigned WIN_BYTES = WIN_ELEMS * sizeof(T);
    __shared__ __align__(128) T s_win[WT_STAGES][WIN_ELEMS];
    __shared__ __align__(8) uint64_t s_full[WT_STAGES];
    __shared__ WindowDesc s_wd[WT_STAGES];
    __shared__ TileInfo s_tile;

    const int tid = threadIdx.x;
    const bool producer = tid >= WT_CONSUMERS;
    const bool issuer = tid == WT_CONSUMERS;
    const T* __restrict__ vol = static_cast<const T*>(args.vol);
    const uint32_t pitch = args.pitch, slice = args.slice_lo;

    if (tid == 0) {
        for (int s = 0; s < WT_STAGES; ++s) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    unsigned g = 0;                         // window sequence number, identical in every thread
    unsigned int n_smem = 0, n_fallback = 0;   // per-thread totals (STATS builds)
    IndexMap im; im.init(fc);
    const int lx_pix = tid & (WT_TILE - 1), ly_pix = (tid >> 4) & (WT_TILE - 1);

    for (;;) {
        // ---- next tile ----
        if (tid == 0) {
            const unsigned t = atomicAdd(args.tile_counter, 1u);
            s_tile.tile = t < (unsigned)(args.tiles_x * args.tiles_y) ? (int)t : -1;
            s_tile.any_hit = 0; s_tile.zs_min = 0x7fffffff; s_tile.zs_max = -0x7fffffff;
        }
        __syncthreads();
        const int tile = s_tile.tile;
        if (tile < 0) break;
        const int tile_x = tile % args.tiles_x, tile_y = tile / args.tiles_x;
        const int px = tile_x * WT_TILE + lx_pix;
        const int lrow = tile_y * WT_TILE + ly_pix;
        const int py = owned_row_to_global(fc, lrow);
        const bool valid = !producer && px < fc.W && lrow < args.local_rows && py < fc.H;

        // ---- ray setup (consumers) / tile geometry (issuer) ----
        RayState rs;
        rs.C = 0.0f; rs.A = 0.0f; rs.iter = 0; rs.active = 0; rs.n_smem = 0; rs.n_fallback = 0;
        CornerRays cr;
        if (valid) {
            bool hit;
            setup_ray_ool(fc, px, py, rs.pos, rs.dstep, hit);
            if (hit) {
                rs.active = 1;
                const int jz0 = __float2int_rd(im.az * rs.pos[2] + im.bz);     // approximate first base index
                atomicMin(&s_tile.zs_min, jz0);
                atomicMax(&s_tile.zs_max, jz0);
                s_tile.any_hit = 1;
            }
        }
        if (issuer) {
            int zm, sg;
            tile_geometry(fc, im, tile_x, tile_y, args.local_rows, cr, zm, sg);
            s_tile.zmajor = zm; s_tile.sgn = sg;
        }
        __syncthreads();
        const bool any_hit = s_tile.any_hit != 0;
        const bool zmajor = s_tile.zmajor != 0;
        const int sgn = s_tile.sgn;

        if (any_hit && !zmajor) {
            // ---- tile not suited to z windows: plain global march ----
            if (rs.active) march_global_tile<T, TCDIV, WIN>(fc, vol, pitch, slice, rs.pos, rs.dstep, rs.C, rs.A);
        } else if (any_hit) {
            // ---- windowed march ----
            const int zs = sgn > 0 ? s_tile.zs_min - 1 : s_tile.zs_max + 1;     // one slice of slack
            int w = 0;
            if (issuer) {
                WindowDesc wd;
                plan_window<T>(cr, sgn > 0 ? zs : zs - (BZ - 2), wd);
                s_wd[g & 1] = wd;
                mbar_expect_tx(&s_full[g & 1], WIN_BYTES);
                for (int k = 0; k < BZ; ++k)
                    tma_load_3d(&s_win[g & 1][k * SLICE_ELEMS], &tmap, &s_full[g & 1], wd.ox0 + k * wd.shx, wd.oy0 + k * wd.shy, wd.a + k);
            }
            for (;;) {
                if (issuer) {      // prefetch window w+1 into the other stage
                    const int a_next = sgn > 0 ? zs + (w + 1) * (BZ - 1) : zs - (BZ - 2) - (w + 1) * (BZ - 1);
                    WindowDesc wd;
                    plan_window<T>(cr, a_next, wd);
                    s_wd[(g + 1) & 1] = wd;
                    mbar_expect_tx(&s_full[(g + 1) & 1], WIN_BYTES);
                    for (int k = 0; k < BZ; ++k)
                        tma_load_3d(&s_win[(g + 1) & 1][k * SLICE_ELEMS], &tmap, &s_full[(g + 1) & 1],
                                    wd.ox0 + k * wd.shx, wd.oy0 + k * wd.shy, wd.a + k);
                }
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                if (rs.active)
                    march_window<T, TCDIV, WIN, STATS>(fc, vol, pitch, slice, smem_u32(&s_win[g & 1][0]), s_wd[g & 1].a,
                                                       s_wd[g & 1].ox0, s_wd[g & 1].oy0, s_wd[g & 1].shy, sgn, rs);
                const int more = __syncthreads_or(rs.active);
                ++g; ++w;
                if (!more) break;
            }
            // the prefetch issued in the last iteration is still in flight: drain it so the
            // stage and its barrier phase are free for the next tile
            mbar_wait(&s_full[g & 1], (g >> 1) & 1);
            ++g;
            __syncthreads();
        }

        if (STATS) { n_smem += rs.n_smem; n_fallback += rs.n_fallback; }
        if (valid) {
            const int orow = fc.compact ? lrow : py;
            reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(rs.C, rs.C, rs.C, rs.A);
        }
        __syncthreads();          // s_tile is rewritten by thread 0 at the top of the loop
    }
    if (STATS && args.stats) {
        if (n_smem) atomicAdd(&args.stats[0], (unsigned long long)n_smem);
        if (n_fallback) atomicAdd(&args.stats[1], (unsigned long long)n_fallback);
        if (tid == 0) atomicAdd(&args.stats[2], (unsigned long long)g);
    }
}

// ------------------------------------------------------------------------------ host side

struct WindowedState {
    CUtensorMap tmap;
    bool tmap_valid = false;
    const void* tmap_vol = nullptr;
    unsigned int* d_counter = nullptr;
    unsigned long long* d_stats = nullptr;
    char err[256] = {0};
};

inline const char* windowed_last_error_of(const WindowedState& st) { return st.err; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

inline void windowed_invalidate(WindowedState& st) { st.tmap_valid = false; st.tmap_vol = nullptr; }
inline void windowed_release(WindowedState& st)
{
    if (st.d_counter) cudaFree(st.d_counter);
    if (st.d_stats) cudaFree(st.d_stats);
    st.d_counter = nullptr; st.d_stats = nullptr; st.tmap_valid = false;
}

// frames the windowed kernel covers: the fast-path preconditions (checked by the caller) plus
// trilinear filtering, a padded volume below 2^32 voxels and a tile-aligned row partition
inline bool windowed_supported(const FrameConsts& fc, int bpv, uint64_t padded_voxels)
{
    (void)bpv;
    return fc.filter == VR_FILTER_TRILINEAR && padded_voxels < (1ull << 32) && (fc.world == 1 || fc.tile_rows % WT_TILE == 0);
}

template <typename T>
int windowed_prepare(WindowedState& st, const void* d_vol, uint32_t pitch, int py, int pz)
{
    if (!st.d_counter && cudaMalloc(&st.d_counter, sizeof(unsigned int)) != cudaSuccess) { snprintf(st.err, sizeof st.err, "cudaMalloc(tile counter) failed"); return -1; }
    if (!st.d_stats && cudaMalloc(&st.d_stats, 3 * sizeof(unsigned long long)) != cudaSuccess) { snprintf(st.err, sizeof st.err, "cudaMalloc(stats) failed"); return -1; }
    if (st.tmap_valid && st.tmap_vol == d_vol) return 0;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) { snprintf(st.err, sizeof st.err, "cuTensorMapEncodeTiled entry point not available"); return -1; }
    const cuuint64_t dims[3] = {pitch, (cuuint64_t)py, (cuuint64_t)pz};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(T), (cuuint64_t)pitch * py * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)WinGeom<T>::BX, (cuuint32_t)WinGeom<T>::BY, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&st.tmap, sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                           const_cast<void*>(d_vol), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(st.err, sizeof st.err, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -1; }
    st.tmap_valid = true; st.tmap_vol = d_vol;
    return 0;
}

// Launches the persistent windowed march.  tcdiv is DIV_RECIP_EXACT or DIV_MARKSTEIN, win a WinMode.
template <typename T>
int launch_windowed_t(WindowedState& st, const FrameConsts& fc, const void* d_vol, uint32_t pitch, uint64_t slice, int dimy, int dimz,
                      float* d_out, int local_rows, int sm_count, int tcdiv, int win, cudaStream_t s, bool want_stats,
                      int ctas_per_sm = 0)
{
    if (windowed_prepare<T>(st, d_vol, pitch, dimy + 2, dimz + 2) != 0) return -1;
    WindowedArgs a{};
    a.vol = d_vol; a.pitch = pitch; a.slice_lo = (uint32_t)slice; a.out = d_out; a.local_rows = local_rows;
    a.tiles_x = (fc.W + WT_TILE - 1) / WT_TILE; a.tiles_y = (local_rows + WT_TILE - 1) / WT_TILE;
    a.tile_counter = st.d_counter; a.stats = want_stats ? st.d_stats : nullptr;
    cudaMemsetAsync(st.d_counter, 0, sizeof(unsigned int), s);
    if (want_stats) cudaMemsetAsync(st.d_stats, 0, 3 * sizeof(unsigned long long), s);
    void (*kern)(const FrameConsts, const WindowedArgs, const CUtensorMap) = nullptr;
    if (want_stats) {
        if (tcdiv == DIV_RECIP_EXACT) kern = win == WIN_COVERS0 ? march_windowed_kernel<T, DIV_RECIP_EXACT, WIN_COVERS0, true> : march_windowed_kernel<T, DIV_RECIP_EXACT, WIN_CLAMP, true>;
        else kern = win == WIN_COVERS0 ? march_windowed_kernel<T, DIV_MARKSTEIN, WIN_COVERS0, true> : march_windowed_kernel<T, DIV_MARKSTEIN, WIN_CLAMP, true>;
    } else {
        if (tcdiv == DIV_RECIP_EXACT) kern = win == WIN_COVERS0 ? march_windowed_kernel<T, DIV_RECIP_EXACT, WIN_COVERS0, false> : march_windowed_kernel<T, DIV_RECIP_EXACT, WIN_CLAMP, false>;
        else kern = win == WIN_COVERS0 ? march_windowed_kernel<T, DIV_MARKSTEIN, WIN_COVERS0, false> : march_windowed_kernel<T, DIV_MARKSTEIN, WIN_CLAMP, false>;
    }
    int per_sm = ctas_per_sm;
    if (per_sm <= 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WT_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const int total_tiles = a.tiles_x * a.tiles_y;
    int grid = sm_count * per_sm;
    if (grid > total_tiles) grid = total_tiles;
    kern<<<grid, WT_THREADS, 0, s>>>(fc, a, st.tmap);
    if (cudaGetLastError() != cudaSuccess) { snprintf(st.err, sizeof st.err, "march_windowed_kernel launch failed"); return -1; }
    return 0;
}

}  // namespace vr
