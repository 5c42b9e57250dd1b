// kernels_aux.cuh -- HBM-bound helper kernels either side of the march:
// volume ingest (edge-replicated padding, min/max, histogram), synthetic volume generation,
// divisor verification, tile assembly after the multi-GPU gather, RGB8 read-back, popcount.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "march_device.cuh"

namespace vr {

// ---- ingest: linear x-fastest volume -> edge-replicated padded volume --------------------
// One thread per 4 padded voxels along x; grid-stride over rows.  Replaces what the GL driver
// does inside glTexImage3D + GL_CLAMP_TO_EDGE (RendererCore.cpp:408-419).
template <typename T>
__global__ void pad_volume_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                  int nx, int ny, int nz, uint32_t pitch)
{
    const uint64_t rows = (uint64_t)(ny + 2) * (uint64_t)(nz + 2);
    for (uint64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int jz = (int)(row / (uint64_t)(ny + 2));
        const int jy = (int)(row - (uint64_t)jz * (uint64_t)(ny + 2));
        const int y = min(max(jy - 1, 0), ny - 1), z = min(max(jz - 1, 0), nz - 1);
        const T* s = src + ((uint64_t)z * ny + y) * (uint64_t)nx;
        T* d = dst + row * (uint64_t)pitch;
        for (uint32_t jx = threadIdx.x; jx < pitch; jx += blockDim.x) {
            const int x = min(max((int)jx - 1, 0), nx - 1);
            d[jx] = s[x];
        }
    }
}

// ---- ingest: pair-packed variant of the padded volume: word x = (voxel x, voxel x+1) ------------
template <typename T, typename W>
__global__ void pad_pairs_kernel(const T* __restrict__ src, W* __restrict__ dst, int nx, int ny, int nz, uint32_t pitch)
{
    const uint64_t rows = (uint64_t)(ny + 2) * (uint64_t)(nz + 2);
    for (uint64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int jz = (int)(row / (uint64_t)(ny + 2));
        const int jy = (int)(row - (uint64_t)jz * (uint64_t)(ny + 2));
        const int y = min(max(jy - 1, 0), ny - 1), z = min(max(jz - 1, 0), nz - 1);
        const T* s = src + ((uint64_t)z * ny + y) * (uint64_t)nx;
        W* d = dst + row * (uint64_t)pitch;
        for (uint32_t jx = threadIdx.x; jx < pitch; jx += blockDim.x) {
            const int x0 = min(max((int)jx - 1, 0), nx - 1), x1 = min(max((int)jx, 0), nx - 1);
            d[jx] = (W)((W)s[x0] | ((W)s[x1] << (8 * sizeof(T))));
        }
    }
}

// ---- ingest: z-pair words for the texpair kernel ------------------------------------------------
// Layers [l0, l0+nl) of the z-pair array: word (x, y, L) = v(x, y, max(L-1,0)) | v(x, y, min(L,nz-1)) << bits.
// L runs over [0, nz]: the layer iz+1 of a sample with iz = floor(fz) in [-1, nz-1] holds both z slices
// of its trilinear footprint, GL_CLAMP_TO_EDGE in z already applied (RendererCore.cpp:413).
template <typename T, typename W>
__global__ void zpair_pack_kernel(const T* __restrict__ src, W* __restrict__ dst, int nx, int ny, int nz, int l0, int nl)
{
    const uint64_t per_layer = (uint64_t)nx * ny, n = per_layer * (uint64_t)nl;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int L = l0 + (int)(i / per_layer);
        const uint64_t xy = i % per_layer;
        const int za = max(L - 1, 0), zb = min(L, nz - 1);
        dst[i] = (W)((W)src[(uint64_t)za * per_layer + xy] | ((W)src[(uint64_t)zb * per_layer + xy] << (8 * sizeof(T))));
    }
}

// Linear copy of the same z-pair words for the LSU stage of the hybrid lab kernels, derived from the
// padded volume (same pitch): word i = padded[i] | padded[i + slice] << bits, i over (nz+1) slices.
template <typename T, typename W>
__global__ void zpair_from_padded_kernel(const T* __restrict__ padded, W* __restrict__ dst, uint64_t slice, uint64_t nwords)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = (W)((W)padded[i] | ((W)padded[i + slice] << (8 * sizeof(T))));
}

// ---- ingest: min/max scan (RendererCore.cpp:362-379) --------------------------------------
// HBM-bound read of the whole volume: 16-byte loads (8 u16 / 16 u8 per thread and request), four
// requests in flight per thread, scalar head/tail for unaligned ends.
template <typename T>
__device__ __forceinline__ void minmax_word(uint32_t w, unsigned int& lo, unsigned int& hi)
{
    if (sizeof(T) == 2) {
        const unsigned int a = w & 0xffffu, b = w >> 16;
        lo = min(lo, min(a, b)); hi = max(hi, max(a, b));
    } else {
        const unsigned int a = w & 0xffu, b = (w >> 8) & 0xffu, c = (w >> 16) & 0xffu, d = w >> 24;
        lo = min(min(lo, min(a, b)), min(c, d)); hi = max(max(hi, max(a, b)), max(c, d));
    }
}

template <typename T>
__global__ void minmax_kernel(const T* __restrict__ src, uint64_t n, unsigned int* out_min, unsigned int* out_max)
{
    unsigned int lo = 0xffffffffu, hi = 0u;
    constexpr uint64_t PER = 16 / sizeof(T);
    const uint64_t head = min(n, (uint64_t)(((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) / sizeof(T)));
    const uint64_t nvec = (n - head) / PER;
    const uint4* __restrict__ v = reinterpret_cast<const uint4*>(src + head);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 a = __ldg(v + i), b = __ldg(v + i + stride), c = __ldg(v + i + 2 * stride), d = __ldg(v + i + 3 * stride);
        minmax_word<T>(a.x, lo, hi); minmax_word<T>(a.y, lo, hi); minmax_word<T>(a.z, lo, hi); minmax_word<T>(a.w, lo, hi);
        minmax_word<T>(b.x, lo, hi); minmax_word<T>(b.y, lo, hi); minmax_word<T>(b.z, lo, hi); minmax_word<T>(b.w, lo, hi);
        minmax_word<T>(c.x, lo, hi); minmax_word<T>(c.y, lo, hi); minmax_word<T>(c.z, lo, hi); minmax_word<T>(c.w, lo, hi);
        minmax_word<T>(d.x, lo, hi); minmax_word<T>(d.y, lo, hi); minmax_word<T>(d.z, lo, hi); minmax_word<T>(d.w, lo, hi);
    }
    for (; i < nvec; i += stride) {
        const uint4 a = __ldg(v + i);
        minmax_word<T>(a.x, lo, hi); minmax_word<T>(a.y, lo, hi); minmax_word<T>(a.z, lo, hi); minmax_word<T>(a.w, lo, hi);
    }
    // scalar head and tail
    for (uint64_t k = tid; k < head; k += stride) { const unsigned int x = src[k]; lo = min(lo, x); hi = max(hi, x); }
    for (uint64_t k = head + nvec * PER + tid; k < n; k += stride) { const unsigned int x = src[k]; lo = min(lo, x); hi = max(hi, x); }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(out_min, lo); atomicMax(out_max, hi); }
}

// ---- ingest: 256-bin histogram (RendererCore.cpp:386-398) ---------------------------------
// 8-bit: bin = value; 16-bit: bin = round(value * 255.0f / max_dataset_val) as the reference computes
// it (float multiply, IEEE divide, round half away, assignment to uint16_t :394); bin 0 and bins > 255
// are skipped.  The 16-bit mapping depends on the value only, so it is tabulated once per upload
// (65536 entries, 0 = skip) and the scan itself is an HBM-bound read: 16-byte loads, the table in
// shared memory, one 256-bin counter set per warp, runs of equal bins merged before the atomic.
__global__ void histogram_lut_kernel(float max_dataset_val, uint8_t* __restrict__ lut)
{
    const unsigned int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 65536u) return;
    const float scaled = roundf(fdiv(fmul((float)v, 255.0f), max_dataset_val));
    const unsigned int b = (unsigned int)scaled & 0xffffu;      // assignment to uint16_t, RendererCore.cpp:394
    lut[v] = (b == 0 || b > 255) ? 0 : (uint8_t)b;
}

constexpr int HIST_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(HIST_THREADS)
histogram_kernel(const T* __restrict__ src, uint64_t n, const uint8_t* __restrict__ lut, unsigned long long* out_bins)
{
    extern __shared__ __align__(16) unsigned char hist_smem[];
    unsigned int* wbins = reinterpret_cast<unsigned int*>(hist_smem);                  // [warps][256]
    uint8_t* slut = hist_smem + (HIST_THREADS / 32) * 256 * sizeof(unsigned int);        // [65536], 16-bit data only
    for (int i = threadIdx.x; i < (HIST_THREADS / 32) * 256; i += blockDim.x) wbins[i] = 0;
    if (sizeof(T) == 2) {
        const uint4* g = reinterpret_cast<const uint4*>(lut);
        uint4* d = reinterpret_cast<uint4*>(slut);
        for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) d[i] = __ldg(g + i);
    }
    __syncthreads();
    unsigned int* mine = wbins + (threadIdx.x >> 5) * 256;
    unsigned int cur = 0, cnt = 0;
    auto add = [&](unsigned int b) {
        if (b == cur) { ++cnt; return; }
        if (cur) atomicAdd(&mine[cur], cnt);
        cur = b; cnt = 1;
    };
    auto add_word = [&](uint32_t w) {
        if (sizeof(T) == 2) { add(slut[w & 0xffffu]); add(slut[w >> 16]); }
        else { add(w & 0xffu); add((w >> 8) & 0xffu); add((w >> 16) & 0xffu); add(w >> 24); }
    };
    constexpr uint64_t PER = 16 / sizeof(T);
    const uint64_t head = min(n, (uint64_t)(((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) / sizeof(T)));
    const uint64_t nvec = (n - head) / PER;
    const uint4* __restrict__ v = reinterpret_cast<const uint4*>(src + head);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = tid;
    for (; i + stride < nvec; i += 2 * stride) {
        const uint4 a = __ldg(v + i), b = __ldg(v + i + stride);
        add_word(a.x); add_word(a.y); add_word(a.z); add_word(a.w);
        add_word(b.x); add_word(b.y); add_word(b.z); add_word(b.w);
    }
    for (; i < nvec; i += stride) {
        const uint4 a = __ldg(v + i);
        add_word(a.x); add_word(a.y); add_word(a.z); add_word(a.w);
    }
    for (uint64_t k = tid; k < head; k += stride) add(sizeof(T) == 2 ? (unsigned int)slut[src[k]] : (unsigned int)src[k]);
    for (uint64_t k = head + nvec * PER + tid; k < n; k += stride) add(sizeof(T) == 2 ? (unsigned int)slut[src[k]] : (unsigned int)src[k]);
    if (cur) atomicAdd(&mine[cur], cnt);
    __syncthreads();
    for (int b = threadIdx.x; b < 256; b += blockDim.x) {
        unsigned int t = 0;
        for (int w = 0; w < HIST_THREADS / 32; ++w) t += wbins[w * 256 + b];
        if (t) atomicAdd(&out_bins[b], (unsigned long long)t);
    }
}

// ---- synthetic volume `mix` --------------------------------------------------
template <typename T>
__global__ void synth_mix_kernel(T* __restrict__ dst, int nx, int ny, int nz, uint32_t vmax,
                                 uint32_t seed, int with_hash)
{
    const uint64_t n = (uint64_t)nx * ny * nz;
    const double TWO_PI = 6.283185307179586476925286766559;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(idx % (uint64_t)nx);
        const uint64_t t = idx / (uint64_t)nx;
        const uint32_t j = (uint32_t)(t % (uint64_t)ny);
        const uint32_t k = (uint32_t)(t / (uint64_t)ny);
        const double px = ((double)i + 0.5) / (double)nx, py = ((double)j + 0.5) / (double)ny,
                     pz = ((double)k + 0.5) / (double)nz;
        const double ddx = px - 0.5, ddy = py - 0.5, ddz = pz - 0.5;
        const double r = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
        const double s1 = (r - 0.30) / 0.04, s2 = (r - 0.15) / 0.03;
        double f = 0.70 * (0.6 * exp(-(s1 * s1)) + 0.4 * exp(-(s2 * s2)))
                 + 0.25 * (0.5 + 0.5 * sin(TWO_PI * (3.0 * px + 0.1)) * sin(TWO_PI * (2.0 * py + 0.2)) * sin(TWO_PI * (5.0 * pz + 0.3)));
        if (with_hash) {
            uint32_t h = (i * 73856093u) ^ (j * 19349663u) ^ (k * 83492791u) ^ seed;
            h *= 2654435761u;
            f += 0.05 * ((double)h / 4294967296.0);
        }
        f = fmin(fmax(f, 0.0), 1.0);
        dst[idx] = (T)(uint32_t)floor((double)vmax * f + 0.5);
    }
}

// ---- Markstein division check: all 2^23 significands against div.rn ------------------------
__global__ void verify_divisor_kernel(float d, float inv, unsigned int* mismatch)
{
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 23)) return;
    const float a = __uint_as_float(0x3F800000u | m);
    const float q = div_by<DIV_MARKSTEIN>(a, d, inv);
    if (__float_as_uint(q) != __float_as_uint(fdiv(a, d))) atomicOr(mismatch, 1u);
}

// ---- multi-GPU: rank-major compact tiles -> full frame (after the NCCL gather) -------------
__global__ void assemble_tiles_kernel(const float4* __restrict__ gathered, float4* __restrict__ frame,
                                      int W, int H, int world, int tile_rows, int rows_per_rank)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    const int tile = y / tile_rows, r_in = y - tile * tile_rows;
    const int rank = tile % world, tile_local = tile / world;
    const int local_row = tile_local * tile_rows + r_in;
    frame[(size_t)y * W + x] = gathered[((size_t)rank * rows_per_rank + local_row) * W + x];
}

// ---- display step: RGBA32F -> RGB8, clamp, round to nearest, optional vertical flip ---------
// (glBlitFramebuffer / glReadPixels(GL_RGB, GL_UNSIGNED_BYTE), RendererCore.cpp:158-171)
__global__ void rgba32f_to_rgb8_kernel(const float4* __restrict__ frame, uint8_t* __restrict__ rgb,
                                       int W, int H, int flip)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W || y >= H) return;
    const float4 p = frame[(size_t)y * W + x];
    const int oy = flip ? (H - 1 - y) : y;
    uint8_t* o = rgb + ((size_t)oy * W + x) * 3;
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float v = c[i];
        v = (v >= 0.0f) ? v : 0.0f;            // NaN -> 0
        v = fminf(v, 1.0f);
        o[i] = (uint8_t)__float2int_rn(fmul(v, 255.0f));
    }
}

// ---- multi-GPU: frame barrier in NVLink peer memory ----------------------------------------------
// Three words behind the owner's frame (same allocation, hence inside every peer's IPC mapping):
// [0] arrivals (cumulative), [1] released frame number, [2] error (a wait gave up).
// signal: every rank, in stream order after its march kernel -- the kernel boundary orders the
// march's peer stores before the system-scope fence + atomic.  wait_*: bounded spins (globaltimer),
// so a dead peer costs `timeout_ns`, never a hung GPU.
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__global__ void peer_signal_kernel(unsigned int* sync)
{
    __threadfence_system();
    atomicAdd_system(&sync[0], 1u);
}
// word `idx` of `sync` >= target (wrap-safe), else error after timeout_ns
__global__ void peer_wait_kernel(unsigned int* sync, int idx, unsigned int target, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(&sync[idx]) - target) < 0) {
        if (global_ns() - t0 > timeout_ns) { atomicExch_system(&sync[2], 1u); break; }
        __nanosleep(64);
    }
    __threadfence_system();
}
__global__ void peer_release_kernel(unsigned int* sync, unsigned int frame_no)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&sync[1]), "r"(frame_no) : "memory");
}

__global__ void popcount_kernel(const unsigned int* __restrict__ bits, uint64_t nwords, unsigned long long* out)
{
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x)
        acc += __popc(bits[i]);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

}  // namespace vr
