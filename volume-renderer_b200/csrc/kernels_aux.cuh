// kernels_aux.cuh -- HBM-bound helper kernels either side of the march:
// volume ingest (edge-replicated padding fused with min/max, histogram, z-pair words, per-cell min/max table),
// synthetic volume generation,
// divisor verification, tile assembly after the multi-GPU gather, RGB8 read-back, popcount.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "march_device.cuh"

namespace vr {

// ---- ingest: linear x-fastest volume -> edge-replicated padded volume, fused with the min/max scan ----
// ONE read of the source: a block per padded row (grid-stride), a thread per padded voxel along x.
// Replaces what the GL driver does inside glTexImage3D + GL_CLAMP_TO_EDGE (RendererCore.cpp:408-419)
// and the reference's single-threaded min/max loop (RendererCore.cpp:362-379); every source voxel
// appears in at least one padded row, so the running min/max over what is copied is the volume's.
template <typename T>
__global__ void pad_minmax_kernel(const T* __restrict__ src, T* __restrict__ dst, int nx, int ny, int nz, uint32_t pitch,
                                  unsigned int* out_min, unsigned int* out_max)
{
    const uint64_t rows = (uint64_t)(ny + 2) * (uint64_t)(nz + 2);
    unsigned int lo = 0xffffffffu, hi = 0u;
    for (uint64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int jz = (int)(row / (uint64_t)(ny + 2));
        const int jy = (int)(row - (uint64_t)jz * (uint64_t)(ny + 2));
        const int y = min(max(jy - 1, 0), ny - 1), z = min(max(jz - 1, 0), nz - 1);
        const T* s = src + ((uint64_t)z * ny + y) * (uint64_t)nx;
        T* d = dst + row * (uint64_t)pitch;
        for (uint32_t jx = threadIdx.x; jx < pitch; jx += blockDim.x) {
            const int x = min(max((int)jx - 1, 0), nx - 1);
            const unsigned int v = s[x];
            d[jx] = (T)v;
            lo = min(lo, v); hi = max(hi, v);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(out_min, lo); atomicMax(out_max, hi); }
}

// ---- ingest, fused: padded copy + per-cell min/max table + dataset min/max in ONE read of the source ---------------
// Fast path of the upload (rows of the source 16-byte aligned: nx a multiple of 16 / sizeof(T)); HBM-bound: 16-byte loads
// and 16-byte stores, no shared-memory staging of the data.  A warp per padded row, 8 rows (one 8-row group of one
// padded slice) per CTA pass:
//  * lane L loads source vector k = 32*i + L; padded vector k is the same data moved up by ONE element (padded index =
//    voxel index + 1), the element shifted in comes from the lane below (__shfl_up; across the 32-vector boundary from
//    lane 31 of the previous pass); vectors past the end of the row replicate the last voxel (GL_CLAMP_TO_EDGE);
//  * the same registers feed the cell table: per aligned group of 8 voxels a SIMD min/max tree (__vminu2 / __vminu4), then
//    one shared-memory atomic pair per group into the CTA's row of cells (all 8 rows of a group lie in the same band of
//    cells because the cell side is a multiple of 8).  Padded index j = voxel + 1 belongs to cell j >> shift and, when j is
//    a multiple of the cell side, also to the cell before it (a trilinear footprint based at j - 1 reaches it): handled
//    along x per group (its last voxel), along y with a second row of cells for the group's first row, along z when the
//    CTA's row of cells is flushed -- one global atomic pair per cell and CTA pass;
//  * rows of the padding (jy or jz = 0, N+1) are copies of real rows and only written.
// `gmin`/`gmax`: 32-bit cell tables (initialised to 0xffffffff / 0), converted to the 16-bit product tables -- together
// with the dataset's min/max, which is the min/max over the cells -- by cell_table_finish_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
pad_cells_kernel(const T* __restrict__ src, T* __restrict__ dst, int nx, int ny, int nz, uint32_t pitch,
                 int shift, int cnx, int cny, unsigned int* __restrict__ gmin, unsigned int* __restrict__ gmax)
{
    constexpr int PER = 16 / (int)sizeof(T), SH = 8 * (int)sizeof(T), G = PER / 8;
    extern __shared__ unsigned int s_cells[];                       // [2 tables][lo | hi][cnx]
    unsigned int* s_lo0 = s_cells, *s_hi0 = s_cells + cnx, *s_lo1 = s_cells + 2 * cnx, *s_hi1 = s_cells + 3 * cnx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, side = 1 << shift;
    const int groups = (ny + 2 + 7) / 8;
    const uint64_t passes = (uint64_t)groups * (uint64_t)(nz + 2);
    const int nvs = nx / PER, nvd = (int)(pitch / PER);
    for (uint64_t it = blockIdx.x; it < passes; it += gridDim.x) {
        const int jz = (int)(it / (uint64_t)groups), jy0 = (int)(it - (uint64_t)jz * groups) * 8, jy = jy0 + warp;
        const bool real_z = jz >= 1 && jz <= nz;
        const bool two_bands = jy0 > 0 && (jy0 & (side - 1)) == 0;                   // the group's first row also belongs to the band before
        if (real_z) {
            for (int i = threadIdx.x; i < 2 * cnx; i += 256) { s_cells[i] = (i < cnx) ? 0xffffffffu : 0u; s_cells[2 * cnx + i] = (i < cnx) ? 0xffffffffu : 0u; }
            __syncthreads();
        }
        if (jy <= ny + 1) {
            const int y = min(max(jy - 1, 0), ny - 1), z = min(max(jz - 1, 0), nz - 1);
            const T* s = src + ((uint64_t)z * ny + y) * (uint64_t)nx;
            const uint4* sv = reinterpret_cast<const uint4*>(s);
            uint4* dv = reinterpret_cast<uint4*>(dst + ((uint64_t)jz * (uint64_t)(ny + 2) + (uint64_t)jy) * (uint64_t)pitch);
            const bool contrib = real_z && jy >= 1 && jy <= ny;
            const bool extra = contrib && warp == 0 && two_bands;
            const unsigned int last = (unsigned int)s[nx - 1];
            unsigned int lastw = last | (last << SH); if (SH == 8) lastw |= lastw << 16;
            unsigned int carry = (unsigned int)s[0];
            uint4 nxt = make_uint4(lastw, lastw, lastw, lastw);
            if (lane < nvs) nxt = __ldg(sv + lane);
            for (int k0 = 0; k0 < nvd; k0 += 32) {
                const int k = k0 + lane;
                const bool has = k < nvs;
                const uint4 own = nxt;                                                 // loaded one pass ahead
                nxt = make_uint4(lastw, lastw, lastw, lastw);
                if (k + 32 < nvs) nxt = __ldg(sv + k + 32);
                const unsigned int own_last = own.w >> (32 - SH);
                unsigned int prev_last = __shfl_up_sync(0xffffffffu, own_last, 1);
                if (lane == 0) prev_last = carry;
                carry = __shfl_sync(0xffffffffu, own_last, 31);
                if (k < nvd) {
                    uint4 o;
                    o.x = (own.x << SH) | prev_last;
                    o.y = __funnelshift_l(own.x, own.y, SH);
                    o.z = __funnelshift_l(own.y, own.z, SH);
                    o.w = __funnelshift_l(own.z, own.w, SH);
                    dv[k] = o;
                }
                if (contrib && has) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        unsigned int lo, hi, tail;
                        if (sizeof(T) == 2) {
                            lo = __vminu2(__vminu2(own.x, own.y), __vminu2(own.z, own.w));
                            hi = __vmaxu2(__vmaxu2(own.x, own.y), __vmaxu2(own.z, own.w));
                            lo = min(lo & 0xffffu, lo >> 16); hi = max(hi & 0xffffu, hi >> 16);
                            tail = own.w >> 16;
                        } else {
                            const unsigned int a = g == 0 ? own.x : own.z, b = g == 0 ? own.y : own.w;
                            lo = __vminu4(a, b); hi = __vmaxu4(a, b);
                            lo = __vminu4(lo, lo >> 16); hi = __vmaxu4(hi, hi >> 16);
                            lo = min(lo & 0xffu, (lo >> 8) & 0xffu); hi = max(hi & 0xffu, (hi >> 8) & 0xffu);
                            tail = b >> 24;
                        }
                        const int j_last = k * PER + g * 8 + 8;                       // padded index of the group's last voxel
                        const int c0 = (j_last - 7) >> shift;
                        atomicMin(&s_lo0[c0], lo); atomicMax(&s_hi0[c0], hi);
                        if (extra) { atomicMin(&s_lo1[c0], lo); atomicMax(&s_hi1[c0], hi); }
                        if ((j_last & (side - 1)) == 0) {                             // ... which also opens the next cell
                            atomicMin(&s_lo0[c0 + 1], tail); atomicMax(&s_hi0[c0 + 1], tail);
                            if (extra) { atomicMin(&s_lo1[c0 + 1], tail); atomicMax(&s_hi1[c0 + 1], tail); }
                        }
                    }
                }
            }
        }
        if (real_z) {
            __syncthreads();
            const int cy = jy0 >> shift, cz = jz >> shift;
            const bool two_slabs = (jz & (side - 1)) == 0;                             // jz >= 1 here
            for (int cx = threadIdx.x; cx < cnx; cx += 256) {
                const unsigned int lo0 = s_lo0[cx], hi0 = s_hi0[cx];
                if (lo0 <= hi0) {
                    const size_t i = ((size_t)cz * cny + cy) * cnx + cx;
                    atomicMin(gmin + i, lo0); atomicMax(gmax + i, hi0);
                    if (two_slabs) { const size_t q = i - (size_t)cny * cnx; atomicMin(gmin + q, lo0); atomicMax(gmax + q, hi0); }
                }
                const unsigned int lo1 = s_lo1[cx], hi1 = s_hi1[cx];
                if (two_bands && lo1 <= hi1) {
                    const size_t i = ((size_t)cz * cny + (cy - 1)) * cnx + cx;
                    atomicMin(gmin + i, lo1); atomicMax(gmax + i, hi1);
                    if (two_slabs) { const size_t q = i - (size_t)cny * cnx; atomicMin(gmin + q, lo1); atomicMax(gmax + q, hi1); }
                }
            }
            __syncthreads();
        }
    }
}

// 32-bit cell tables of pad_cells_kernel -> the product's 16-bit tables, plus the dataset min/max (= over all cells)
__global__ void cell_table_finish_kernel(const unsigned int* __restrict__ gmin, const unsigned int* __restrict__ gmax, uint64_t ncells,
                                         uint16_t* __restrict__ cmin, uint16_t* __restrict__ cmax, unsigned int* out_min, unsigned int* out_max)
{
    unsigned int lo = 0xffffffffu, hi = 0u;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned int a = gmin[i], b = gmax[i];
        cmin[i] = (uint16_t)a; cmax[i] = (uint16_t)b;
        lo = min(lo, a); hi = max(hi, b);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(out_min, lo); atomicMax(out_max, hi); }
}

// ---- z-pair words for the texpair kernel, derived from the padded volume -----------------------------
// Layers [l0, l0+nl) of the z-pair array into a tightly packed staging buffer: word (x, y, L) =
// v(x, y, max(L-1,0)) | v(x, y, min(L,nz-1)) << bits = padded(x+1, y+1, L) | padded(x+1, y+1, L+1) << bits.
// L runs over [0, nz]: layer iz+1 of a sample with iz = floor(fz) in [-1, nz-1] holds both z slices of its
// trilinear footprint, GL_CLAMP_TO_EDGE in z already applied (RendererCore.cpp:413).
template <typename T, typename W>
__global__ void zpair_pack_kernel(const T* __restrict__ padded, W* __restrict__ dst, int nx, int ny, uint32_t pitch,
                                  uint64_t slice, int l0, int nl)
{
    const uint64_t per_layer = (uint64_t)nx * ny, n = per_layer * (uint64_t)nl;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int L = l0 + (int)(i / per_layer);
        const uint64_t xy = i % per_layer;
        const uint32_t y = (uint32_t)(xy / (uint64_t)nx), x = (uint32_t)(xy - (uint64_t)y * nx);
        const uint64_t e = (uint64_t)L * slice + (uint64_t)(y + 1) * pitch + (x + 1);
        dst[i] = (W)((W)padded[e] | ((W)padded[e + slice] << (8 * sizeof(T))));
    }
}

// ---- per-cell min/max table (the brick table of SURVEY 8f-2) and the empty-cell map ------------------
// Cell (cx,cy,cz) of side 2^shift covers, on each axis, the voxel indices [c*2^shift - 1, (c+1)*2^shift - 1]
// clamped to the volume: exactly the voxels a sample whose base index i0 satisfies (i0+1) >> shift == c can
// touch (trilinear footprint {i0, i0+1} after GL_CLAMP_TO_EDGE; the nearest voxel is i0 itself).  One CTA per
// cell, grid-stride; a warp per (y,z) row of the cell, lanes along x.
template <typename T>
__global__ void __launch_bounds__(128)
cell_minmax_kernel(const T* __restrict__ padded, uint32_t pitch, uint64_t slice, int nx, int ny, int nz, int shift,
                   int cnx, int cny, int cnz, uint16_t* __restrict__ cmin, uint16_t* __restrict__ cmax)
{
    __shared__ unsigned int s_lo[4], s_hi[4];
    const int side = 1 << shift, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t ncells = (uint64_t)cnx * cny * cnz;
    for (uint64_t cell = blockIdx.x; cell < ncells; cell += gridDim.x) {
        const int cx = (int)(cell % (uint64_t)cnx), cy = (int)((cell / (uint64_t)cnx) % (uint64_t)cny), cz = (int)(cell / ((uint64_t)cnx * cny));
        // padded index = voxel index + 1; the padding replicates the edge, so the clamp is only needed to stay in bounds
        const int x0 = cx * side, x1 = min((cx + 1) * side, nx + 1);       // padded x range [x0, x1]
        const int y0 = cy * side, y1 = min((cy + 1) * side, ny + 1);
        const int z0 = cz * side, z1 = min((cz + 1) * side, nz + 1);
        unsigned int lo = 0xffffffffu, hi = 0u;
        const int rows_y = y1 - y0 + 1, nrows = rows_y * (z1 - z0 + 1);
        for (int r = warp; r < nrows; r += 4) {
            const int jz = z0 + r / rows_y, jy = y0 + r % rows_y;
            const T* row = padded + (uint64_t)jz * slice + (uint64_t)jy * pitch;
            for (int jx = x0 + lane; jx <= x1; jx += 32) { const unsigned int v = row[jx]; lo = min(lo, v); hi = max(hi, v); }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            cmin[cell] = (uint16_t)min(min(s_lo[0], s_lo[1]), min(s_lo[2], s_lo[3]));
            cmax[cell] = (uint16_t)max(max(s_hi[0], s_hi[1]), max(s_hi[2], s_hi[3]));
        }
        __syncthreads();
    }
}

// bit c of the map = (cell max <= threshold); one warp ballot per 32 cells; counts the empty cells.
// Launch with a whole number of warps covering ceil(ncells / 32) * 32 cells.
__global__ void cell_empty_kernel(const uint16_t* __restrict__ cmax, uint64_t ncells, unsigned int threshold,
                                  uint32_t* __restrict__ bits, unsigned long long* count)
{
    const uint64_t words = (ncells + 31) / 32;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    unsigned int n = 0;
    for (uint64_t w = warp; w < words; w += nwarps) {
        const uint64_t i = w * 32 + lane;
        const bool e = i < ncells && (unsigned int)cmax[i] <= threshold;
        const unsigned int m = __ballot_sync(0xffffffffu, e);
        if (lane == 0) { bits[w] = m; n += __popc(m); }
    }
    if (lane == 0 && n) atomicAdd(count, (unsigned long long)n);
}

// ---- ingest: 256-bin histogram (RendererCore.cpp:386-398) ---------------------------------
// 8-bit: bin = value; 16-bit: bin = round(value * 255.0f / max_dataset_val) as the reference computes
// it (float multiply, IEEE divide, round half away, assignment to uint16_t :394); bin 0 and bins > 255
// are skipped.  The 16-bit mapping depends on the value only, so it is tabulated once per upload
// (65536 entries, 0 = skip; only entries [0, dataset max] are ever read, so only those are staged: 4 KB instead of
// 64 KB for 12-bit data, which triples the resident CTAs) and the scan itself is an HBM-bound read: 16-byte loads, the
// table in shared memory, one 256-bin counter set per warp, runs of equal bins merged before the atomic.
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
histogram_kernel(const T* __restrict__ src, uint64_t n, const uint8_t* __restrict__ lut, int lut_entries, unsigned long long* out_bins)
{
    extern __shared__ __align__(16) unsigned char hist_smem[];
    unsigned int* wbins = reinterpret_cast<unsigned int*>(hist_smem);                  // [warps][256]
    uint8_t* slut = hist_smem + (HIST_THREADS / 32) * 256 * sizeof(unsigned int);        // [lut_entries], 16-bit data only
    for (int i = threadIdx.x; i < (HIST_THREADS / 32) * 256; i += blockDim.x) wbins[i] = 0;
    if (sizeof(T) == 2) {
        const uint4* g = reinterpret_cast<const uint4*>(lut);
        uint4* d = reinterpret_cast<uint4*>(slut);
        for (int i = threadIdx.x; i < lut_entries / 16; i += blockDim.x) d[i] = __ldg(g + i);   // entries [0, dataset max], rounded up to 16
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
// (`host_error`: a word of mapped pinned host memory the next ABI call checks without synchronising)
__global__ void peer_wait_kernel(unsigned int* sync, int idx, unsigned int target, unsigned long long timeout_ns,
                                 volatile unsigned int* host_error)
{
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(&sync[idx]) - target) < 0) {
        if (global_ns() - t0 > timeout_ns) {
            atomicExch_system(&sync[2], 1u);
            if (host_error) { *host_error = 1u + (unsigned int)idx; __threadfence_system(); }
            break;
        }
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

// ---- launch order of the march kernels' CTA tiles: longest rays first ---------------------------------------------
// One CTA.  Pass 1: every tile's cost = the longest chord of its centre and corner rays through the box (computeRay /
// intersectRayAABB, VolumeRenderer.cs:194-238, evaluated approximately -- it only orders work; centre + two opposite
// corners), quantised into 1024
// classes (class 0 = longest), histogram in shared memory.  Pass 2: exclusive scan of the classes.  Pass 3: scatter,
// order[offset[class]++] = tile x | y << 16.  Every tile appears exactly once whatever the atomics' order.
__device__ __forceinline__ float ray_chord(const FrameConsts& fc, float pxc, float pyc)
{
    const float* cam = fc.cam;
    const float aspect = (float)fc.W / (float)fc.H;
    const float x = aspect * (2.0f * pxc / (float)fc.W - 1.0f), y = 2.0f * pyc / (float)fc.H - 1.0f, z = -cam[20];
    const float r0 = rsqrtf(x * x + y * y + z * z);
    const float dx = x * r0, dy = y * r0, dz = z * r0;
    float m[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) m[i] = cam[i] * dx + cam[4 + i] * dy + cam[8 + i] * dz;
    const float r1 = rsqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    float t0 = -3.0e38f, t1 = 3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float inv = 1.0f / (m[i] * r1);
        const float a = (fc.pmin[i] - cam[16 + i]) * inv, b = (fc.pmax[i] - cam[16 + i]) * inv;
        t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    }
    const float len = t1 - fmaxf(t0, 0.0f);
    return len > 0.0f ? len : 0.0f;                                     // NaN -> 0
}

__global__ void __launch_bounds__(1024)
cta_order_kernel(const __grid_constant__ FrameConsts fc, int row0, int row_end, int px_w, int px_h, int gx, int gy,
                 uint32_t* __restrict__ order, uint32_t* __restrict__ scratch)
{
    constexpr int CLASSES = 1024;
    __shared__ unsigned int s_count[CLASSES];
    const int n = gx * gy;
    s_count[threadIdx.x] = 0u;
    __syncthreads();
    const float ex = fc.pmax[0] - fc.pmin[0], ey = fc.pmax[1] - fc.pmin[1], ez = fc.pmax[2] - fc.pmin[2];
    const float scale = (float)(CLASSES - 1) * rsqrtf(ex * ex + ey * ey + ez * ez);     // chord <= diagonal
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int by = i / gx, bx = i - by * gx;
        const int l0 = row0 + by * px_h, l1 = min(l0 + px_h, row_end) - 1;
        const float y0 = (float)min(owned_row_to_global(fc, l0), fc.H - 1) + 0.5f, y1 = (float)min(owned_row_to_global(fc, l1), fc.H - 1) + 0.5f;
        const float x0 = (float)(bx * px_w) + 0.5f, x1 = (float)min(bx * px_w + px_w, fc.W) - 0.5f;
        const float c = fmaxf(fmaxf(ray_chord(fc, x0, y0), ray_chord(fc, x1, y1)), ray_chord(fc, 0.5f * (x0 + x1), 0.5f * (y0 + y1)));
        const int cls = CLASSES - 1 - min(max((int)(c * scale), 0), CLASSES - 1);
        scratch[i] = (uint32_t)cls;
        atomicAdd(&s_count[cls], 1u);
    }
    __syncthreads();
    {                                                                    // exclusive scan, 1024 classes = one per thread
        __shared__ unsigned int s_warp[32];
        const unsigned int v = s_count[threadIdx.x];
        unsigned int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, d); if ((threadIdx.x & 31) >= d) inc += u; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int w = s_warp[threadIdx.x], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, winc, d); if (threadIdx.x >= d) winc += u; }
            s_warp[threadIdx.x] = winc - w;
        }
        __syncthreads();
        s_count[threadIdx.x] = s_warp[threadIdx.x >> 5] + inc - v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int by = i / gx, bx = i - by * gx;
        const unsigned int pos = atomicAdd(&s_count[scratch[i]], 1u);
        order[pos] = (uint32_t)bx | ((uint32_t)by << 16);
    }
}

}  // namespace vr
