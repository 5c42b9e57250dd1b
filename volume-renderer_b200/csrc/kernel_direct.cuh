// kernel_direct.cuh -- one thread per ray, voxels fetched straight from the padded volume in
// HBM through L1/L2 (ld.global.nc).  The simplest correct device path: it is the fallback
// of the windowed kernel, the kernel behind every GENERIC (MIP / TF / view swizzle) frame,
// and the instrumentation pass that counts distinct voxels.
#pragma once

#include "march_device.cuh"

namespace vr {

constexpr int DIRECT_BLOCK_W = 32;   // pixels per CTA in x
constexpr int DIRECT_BLOCK_H = 8;    // pixels per CTA in y; a warp covers an 8x4 pixel patch

struct DirectArgs {
    const void* vol;
    uint32_t pitch;
    uint64_t slice;
    const float* tf_lut;
    float* out;
    int local_rows;                  // rows owned by this rank
    unsigned int* touch_bits;        // COUNT only
    unsigned long long* counters;    // COUNT only: [0] samples, [1] rays hit
};

template <typename T, int FILTER, int TCDIV, bool GENERIC, bool COUNT>
__global__ void __launch_bounds__(DIRECT_BLOCK_W * DIRECT_BLOCK_H)
march_direct_kernel(const __grid_constant__ FrameConsts fc, const __grid_constant__ DirectArgs args)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * DIRECT_BLOCK_W + (warp & 3) * 8 + (lane & 7);
    const int lrow = fc.row0 + blockIdx.y * DIRECT_BLOCK_H + (warp >> 2) * 4 + (lane >> 3);
    if (px >= fc.W || lrow >= args.local_rows) return;
    const int py = owned_row_to_global(fc, lrow);
    if (py >= fc.H) return;

    PaddedVolume<T> vol{static_cast<const T*>(args.vol), args.pitch, args.slice};
    TouchMap tm{args.touch_bits, fc.dim[0], fc.dim[1], fc.dim[2]};

    float C = 0.0f, A = 0.0f;
    unsigned long long nsamples = 0;
    const RaySetup r = setup_ray(fc, px, py);
    if (r.hit) march_ray<T, FILTER, TCDIV, GENERIC, COUNT>(fc, vol, args.tf_lut, r, tm, C, A, nsamples);

    const int orow = fc.compact ? lrow : py;
    reinterpret_cast<float4*>(args.out)[(size_t)orow * fc.W + px] = make_float4(C, C, C, A);

    if (COUNT) {
        if (nsamples) atomicAdd(&args.counters[0], nsamples);
        if (r.hit) atomicAdd(&args.counters[1], 1ull);
    }
}

}  // namespace vr
