// lab_aux.cuh -- ingest helpers of the round-1 development kernels (edge-replicated padding, x-pair words, z-pair
// words from a plain source); the product's ingest lives in volume-renderer_b200/csrc/kernels_aux.cuh.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace vr {
namespace lab {

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
__global__ void zpair_pack_from_source_kernel(const T* __restrict__ src, W* __restrict__ dst, int nx, int ny, int nz, int l0, int nl)
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

}  // namespace lab
}  // namespace vr
