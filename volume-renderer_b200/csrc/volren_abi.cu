// volren_abi.cu -- context management and the C-ABI of libvolren_b200.so (include/volren_b200.h).
//
// Host code in this file evaluates the per-frame constants (frame.h) and must be compiled
// with -Xcompiler -ffp-contract=off.  There is no CPU fallback anywhere in this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "volren_b200.h"
#include "frame.h"
#include "march_device.cuh"
#include "kernel_direct.cuh"
#include "kernel_fast.cuh"
#include "kernel_windowed.cuh"
#include "kernels_aux.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    cudaGetLastError();   // clear sticky-less errors
    return fail(e == cudaErrorMemoryAllocation ? VR_ERR_OOM : VR_ERR_CUDA, buf);
}

#define VR_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);           \
    } while (0)

inline uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// the frame allocation carries the peer-barrier words behind the pixels (kernels_aux.cuh)
constexpr size_t FRAME_SYNC_BYTES = 256;
inline size_t frame_alloc_bytes(int w, int h) { return (size_t)w * h * 4 * sizeof(float) + FRAME_SYNC_BYTES; }

}  // namespace

struct vr_context {
    int device = 0;
    int W = 0, H = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float* d_frame = nullptr;            // W*H*4
    uint8_t* d_rgb8 = nullptr;           // W*H*3
    // volume (padded, edge replicated)
    void* d_vol = nullptr;
    uint64_t vol_bytes = 0;
    int32_t dim[3] = {0, 0, 0};
    int bpv = 0;
    uint32_t pitch = 0;                  // elements
    uint64_t slice = 0;                  // elements
    // second copy for the texture-gather kernel: layered 2-D array (layer = z), point sampling
    cudaArray_t d_arr = nullptr;
    cudaTextureObject_t tex = 0;
    // third copy for the z-pair gather kernel: layered 2-D array of (z, z+1) words, Nz+1 layers
    cudaArray_t d_arr2 = nullptr;
    cudaTextureObject_t tex2 = 0;
    // ... and the same words as a linear edge-replicated array for the LSU stage of the hybrid kernel
    void* d_zlin = nullptr;
    uint32_t zpitch = 0;                 // words
    uint64_t zslice = 0;                 // words
    float voxel_size[3] = {1.f, 1.f, 1.f};
    vr_volume_stats stats{};
    bool have_stats = false;
    // state
    float cam[21];
    bool have_cam = false;
    vr_params params;
    float* d_lut = nullptr;
    bool lut_fast_ok = false;            // every LUT entry finite and >= +0
    int rank = 0, world = 1, tile_rows = 8;
    // Markstein verification cache: divisor bits -> ok
    std::map<uint32_t, bool> div_ok;
    unsigned int* d_flag = nullptr;
    // windowed kernel scratch
    vr::WindowedState win;
    // vr_render to a host buffer: row bands on their own streams so a band's device->host copy
    // overlaps the march of the following bands
    static constexpr int BANDS = 4;
    cudaStream_t band_stream[BANDS] = {};
    cudaEvent_t band_kdone[BANDS] = {}, band_cdone[BANDS] = {};
    bool bands_ready = false;
};

namespace {

int owned_rows_of(int H, int rank, int world, int tile_rows)
{
    const int tiles = (H + tile_rows - 1) / tile_rows;
    int rows = 0;
    for (int t = rank; t < tiles; t += world) {
        const int y0 = t * tile_rows;
        rows += std::min(tile_rows, H - y0);
    }
    return rows;
}

// rows of the compact image: every owned tile padded to tile_rows (keeps the gather regular)
int compact_rows_of(int H, int rank, int world, int tile_rows)
{
    const int tiles = (H + tile_rows - 1) / tile_rows;
    const int owned_tiles = (tiles - rank + world - 1) / world;
    (void)rank;
    return owned_tiles * tile_rows;
}

int max_compact_rows(int H, int world, int tile_rows)
{
    return compact_rows_of(H, 0, world, tile_rows);
}

int verify_divisor(vr_context* c, float d, bool* ok)
{
    uint32_t bits;
    std::memcpy(&bits, &d, 4);
    auto it = c->div_ok.find(bits);
    if (it != c->div_ok.end()) { *ok = it->second; return VR_OK; }
    if (!(d > 0.0f) || std::isinf(d)) { c->div_ok[bits] = false; *ok = false; return VR_OK; }
    VR_CUDA(cudaMemsetAsync(c->d_flag, 0, sizeof(unsigned int), c->stream));
    vr::verify_divisor_kernel<<<(1u << 23) / 256, 256, 0, c->stream>>>(d, 1.0f / d, c->d_flag);
    VR_CUDA(cudaGetLastError());
    unsigned int flag = 1;
    VR_CUDA(cudaMemcpyAsync(&flag, c->d_flag, sizeof flag, cudaMemcpyDeviceToHost, c->stream));
    VR_CUDA(cudaStreamSynchronize(c->stream));
    c->div_ok[bits] = (flag == 0);
    *ok = (flag == 0);
    return VR_OK;
}

struct LaunchPlan {
    vr::FrameConsts fc;
    bool generic;
    int tcdiv;
    int local_rows;
};

int make_plan(vr_context* c, int compact, LaunchPlan* plan)
{
    if (!c->d_vol) return fail(VR_ERR_NO_VOLUME, "render: no volume uploaded");
    if (!c->have_cam) return fail(VR_ERR_INVALID, "render: no camera set");
    vr::FrameConsts& fc = plan->fc;
    std::memset(&fc, 0, sizeof fc);
    vr::compute_frame_consts(fc, c->W, c->H, c->dim, c->voxel_size, c->cam, c->params);
    fc.rank = c->rank; fc.world = c->world; fc.tile_rows = c->tile_rows; fc.compact = compact;
    plan->local_rows = compact_rows_of(c->H, c->rank, c->world, c->tile_rows);

    const vr_params& p = c->params;
    // the transfer-function LUT alone does not need the generic loop: the pipelined gather kernel has a TF form
    const bool tf_fast = p.use_tf != 0 && c->lut_fast_ok && c->tex2 != 0 && p.filter == VR_FILTER_TRILINEAR &&
                         (p.kernel == VR_KERNEL_AUTO || p.kernel == VR_KERNEL_TEXPAIR_PIPE);
    const bool pipe_selectable = c->tex2 != 0 && p.filter == VR_FILTER_TRILINEAR &&
                                 (p.kernel == VR_KERNEL_AUTO || p.kernel == VR_KERNEL_TEXPAIR_PIPE);
    const bool mip_fast = p.is_mip == 1 && p.use_tf == 0 && p.view_top != 1 && p.view_bottom != 1 && pipe_selectable;   // ... a MIP form
    const bool swizzled = p.view_top == 1 || p.view_bottom == 1;
    const bool view_fast = swizzled && p.is_mip != 1 && p.use_tf == 0 && pipe_selectable;   // ... and view_top / view_bottom forms
    bool generic = (p.is_mip == 1 && !mip_fast) || (p.use_tf != 0 && !(tf_fast && p.is_mip != 1 && !swizzled)) ||
                   (swizzled && !view_fast) ||
                   fc.opacity_correction || !(p.max_val > p.min_val);
    if (!generic) {
        bool ok = false;
        int rc = verify_divisor(c, fc.frange, &ok);
        if (rc != VR_OK) return rc;
        if (!ok) generic = true;
    }
    int tcdiv = fc.tc_div_mode;
    if (tcdiv == vr::DIV_MARKSTEIN) {
        for (int i = 0; i < 3; ++i) {
            if (vr::is_pow2_float(fc.denom[i])) continue;
            bool ok = false;
            int rc = verify_divisor(c, fc.denom[i], &ok);
            if (rc != VR_OK) return rc;
            if (!ok) tcdiv = vr::DIV_IEEE;
        }
    }
    if (generic) tcdiv = vr::DIV_IEEE;
    fc.tc_div_mode = tcdiv;
    plan->generic = generic;
    plan->tcdiv = tcdiv;
    return VR_OK;
}

template <typename T, bool COUNT>
int launch_direct_t(vr_context* c, const LaunchPlan& plan, const vr::DirectArgs& args, cudaStream_t s)
{
    using namespace vr;
    const dim3 block(DIRECT_BLOCK_W * DIRECT_BLOCK_H);
    const dim3 grid((c->W + DIRECT_BLOCK_W - 1) / DIRECT_BLOCK_W,
                    (plan.local_rows + DIRECT_BLOCK_H - 1) / DIRECT_BLOCK_H);
    const FrameConsts& fc = plan.fc;
    if (COUNT || plan.generic) {
        march_direct_kernel<T, VR_FILTER_NEAREST, DIV_IEEE, true, COUNT><<<grid, block, 0, s>>>(fc, args);
    } else if (fc.filter == VR_FILTER_NEAREST) {
        switch (plan.tcdiv) {
            case DIV_RECIP_EXACT: march_direct_kernel<T, VR_FILTER_NEAREST, DIV_RECIP_EXACT, false, false><<<grid, block, 0, s>>>(fc, args); break;
            case DIV_MARKSTEIN:   march_direct_kernel<T, VR_FILTER_NEAREST, DIV_MARKSTEIN, false, false><<<grid, block, 0, s>>>(fc, args); break;
            default:              march_direct_kernel<T, VR_FILTER_NEAREST, DIV_IEEE, false, false><<<grid, block, 0, s>>>(fc, args); break;
        }
    } else {
        switch (plan.tcdiv) {
            case DIV_RECIP_EXACT: march_direct_kernel<T, VR_FILTER_TRILINEAR, DIV_RECIP_EXACT, false, false><<<grid, block, 0, s>>>(fc, args); break;
            case DIV_MARKSTEIN:   march_direct_kernel<T, VR_FILTER_TRILINEAR, DIV_MARKSTEIN, false, false><<<grid, block, 0, s>>>(fc, args); break;
            default:              march_direct_kernel<T, VR_FILTER_TRILINEAR, DIV_IEEE, false, false><<<grid, block, 0, s>>>(fc, args); break;
        }
    }
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

int launch_direct(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s)
{
    vr::DirectArgs args{};
    args.vol = c->d_vol; args.pitch = c->pitch; args.slice = c->slice;
    args.tf_lut = c->d_lut; args.out = d_out; args.local_rows = plan.local_rows;
    return c->bpv == 1 ? launch_direct_t<uint8_t, false>(c, plan, args, s)
                       : launch_direct_t<uint16_t, false>(c, plan, args, s);
}

template <typename T, int WIN>
void launch_packed_tw(const vr::FrameConsts& fc, const vr::FastArgs& a, dim3 grid, cudaStream_t s, bool unit, bool recip, bool nocap)
{
    using namespace vr;
    const dim3 block(256);
    if (unit && nocap)        march_packed_kernel<T, DIV_RECIP_EXACT, WIN, true, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip && nocap)  march_packed_kernel<T, DIV_RECIP_EXACT, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip)           march_packed_kernel<T, DIV_RECIP_EXACT, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
    else if (nocap)           march_packed_kernel<T, DIV_MARKSTEIN, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else                      march_packed_kernel<T, DIV_MARKSTEIN, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
}

template <typename T, int WIN>
void launch_tex_tw(const vr::FrameConsts& fc, const vr::TexArgs& a, dim3 grid, cudaStream_t s, bool unit, bool recip, bool nocap)
{
    using namespace vr;
    const dim3 block(256);
    if (unit && nocap)        march_texgather_kernel<T, DIV_RECIP_EXACT, WIN, true, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip && nocap)  march_texgather_kernel<T, DIV_RECIP_EXACT, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip)           march_texgather_kernel<T, DIV_RECIP_EXACT, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
    else if (nocap)           march_texgather_kernel<T, DIV_MARKSTEIN, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else                      march_texgather_kernel<T, DIV_MARKSTEIN, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
}

// loop-shape flags shared by the packed kernels
void packed_flags(const vr_context* c, const LaunchPlan& plan, bool* unit, bool* recip, bool* nocap)
{
    const vr::FrameConsts& fc = plan.fc;
    *recip = plan.tcdiv == vr::DIV_RECIP_EXACT;
    *unit = *recip && fc.denom[0] == 1.0f && fc.denom[1] == 1.0f && fc.denom[2] == 1.0f;
    // VolumeRenderer.cs:115 caps the loop at 10000 iterations; a ray cannot take more than
    // |box diagonal| / step + 2 = |vol_size| / step_scale + 2 samples
    const double nmax = std::sqrt((double)c->dim[0] * c->dim[0] + (double)c->dim[1] * c->dim[1] + (double)c->dim[2] * c->dim[2]) /
                        (double)fc.step_scale + 4.0;
    *nocap = nmax < 10000.0;
}

int launch_texgather(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win)
{
    vr::TexArgs a{};
    a.tex = c->tex; a.out = d_out; a.local_rows = plan.local_rows;
    const dim3 grid((c->W + 31) / 32, (plan.local_rows + 7) / 8);
    bool unit, recip, nocap;
    packed_flags(c, plan, &unit, &recip, &nocap);
    // the element type only matters for the array's channel format; uint16_t instantiation serves both
    if (win == vr::WIN_COVERS0) launch_tex_tw<uint16_t, vr::WIN_COVERS0>(plan.fc, a, grid, s, unit, recip, nocap);
    else                        launch_tex_tw<uint16_t, vr::WIN_CLAMP>(plan.fc, a, grid, s, unit, recip, nocap);
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

template <int WIN>
void launch_nearest_tex_w(const vr::FrameConsts& fc, const vr::TexArgs& a, dim3 grid, cudaStream_t s, bool unit, bool recip, bool nocap)
{
    using namespace vr;
    const dim3 block(256);
    if (unit && nocap)        march_nearest_tex_kernel<DIV_RECIP_EXACT, WIN, true, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip && nocap)  march_nearest_tex_kernel<DIV_RECIP_EXACT, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip)           march_nearest_tex_kernel<DIV_RECIP_EXACT, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
    else if (nocap)           march_nearest_tex_kernel<DIV_MARKSTEIN, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else                      march_nearest_tex_kernel<DIV_MARKSTEIN, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
}

// nearest filter: integer-coordinate texel loads from the source-type layered array, software pipelined
int launch_nearest_tex(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win)
{
    vr::TexArgs a{};
    a.tex = c->tex; a.out = d_out; a.local_rows = plan.local_rows;
    const dim3 grid((c->W + 31) / 32, (plan.local_rows + 7) / 8);
    bool unit, recip, nocap;
    packed_flags(c, plan, &unit, &recip, &nocap);
    if (win == vr::WIN_COVERS0) launch_nearest_tex_w<vr::WIN_COVERS0>(plan.fc, a, grid, s, unit, recip, nocap);
    else                        launch_nearest_tex_w<vr::WIN_CLAMP>(plan.fc, a, grid, s, unit, recip, nocap);
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

template <typename T, int WIN>
void launch_texpair_tw(const vr::FrameConsts& fc, const vr::TexArgs& a, dim3 grid, cudaStream_t s, bool unit, bool recip, bool nocap)
{
    using namespace vr;
    const dim3 block(256);
    if (unit && nocap)        march_texpair_kernel<T, DIV_RECIP_EXACT, WIN, true, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip && nocap)  march_texpair_kernel<T, DIV_RECIP_EXACT, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else if (recip)           march_texpair_kernel<T, DIV_RECIP_EXACT, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
    else if (nocap)           march_texpair_kernel<T, DIV_MARKSTEIN, WIN, false, true><<<grid, block, 0, s>>>(fc, a);
    else                      march_texpair_kernel<T, DIV_MARKSTEIN, WIN, false, false><<<grid, block, 0, s>>>(fc, a);
}

int launch_texpair(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win)
{
    vr::TexArgs a{};
    a.tex = c->tex2; a.out = d_out; a.local_rows = plan.local_rows;
    const dim3 grid((c->W + 31) / 32, (plan.local_rows + 7) / 8);
    bool unit, recip, nocap;
    packed_flags(c, plan, &unit, &recip, &nocap);
    if (c->bpv == 2) {
        if (win == vr::WIN_COVERS0) launch_texpair_tw<uint16_t, vr::WIN_COVERS0>(plan.fc, a, grid, s, unit, recip, nocap);
        else                        launch_texpair_tw<uint16_t, vr::WIN_CLAMP>(plan.fc, a, grid, s, unit, recip, nocap);
    } else {
        if (win == vr::WIN_COVERS0) launch_texpair_tw<uint8_t, vr::WIN_COVERS0>(plan.fc, a, grid, s, unit, recip, nocap);
        else                        launch_texpair_tw<uint8_t, vr::WIN_CLAMP>(plan.fc, a, grid, s, unit, recip, nocap);
    }
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

template <typename T, int WIN, int MINB>
void launch_texpair2_tw(const vr::FrameConsts& fc, const vr::TexArgs& a, dim3 grid, cudaStream_t s, bool unit, bool recip, bool nocap)
{
    using namespace vr;
    const dim3 block(256);
    if (unit && nocap)        march_texpair2_kernel<T, DIV_RECIP_EXACT, WIN, true, true, MINB><<<grid, block, 0, s>>>(fc, a);
    else if (recip && nocap)  march_texpair2_kernel<T, DIV_RECIP_EXACT, WIN, false, true, MINB><<<grid, block, 0, s>>>(fc, a);
    else if (recip)           march_texpair2_kernel<T, DIV_RECIP_EXACT, WIN, false, false, MINB><<<grid, block, 0, s>>>(fc, a);
    else if (nocap)           march_texpair2_kernel<T, DIV_MARKSTEIN, WIN, false, true, MINB><<<grid, block, 0, s>>>(fc, a);
    else                      march_texpair2_kernel<T, DIV_MARKSTEIN, WIN, false, false, MINB><<<grid, block, 0, s>>>(fc, a);
}

template <typename T, int WIN, int FA, int FB, int MODE>
void launch_texpair_pipe_tw(const vr::FrameConsts& fc, const vr::TexArgs& a, dim3 grid, cudaStream_t s, bool unit, bool recip, bool nocap)
{
    using namespace vr;
    const dim3 block(256);
    if (unit && nocap)        march_texpair_pipe_kernel<T, DIV_RECIP_EXACT, WIN, true, true, 6, FA, FB, MODE><<<grid, block, 0, s>>>(fc, a);
    else if (recip && nocap)  march_texpair_pipe_kernel<T, DIV_RECIP_EXACT, WIN, false, true, 6, FA, FB, MODE><<<grid, block, 0, s>>>(fc, a);
    else if (recip)           march_texpair_pipe_kernel<T, DIV_RECIP_EXACT, WIN, false, false, 6, FA, FB, MODE><<<grid, block, 0, s>>>(fc, a);
    else if (nocap)           march_texpair_pipe_kernel<T, DIV_MARKSTEIN, WIN, false, true, 6, FA, FB, MODE><<<grid, block, 0, s>>>(fc, a);
    else                      march_texpair_pipe_kernel<T, DIV_MARKSTEIN, WIN, false, false, 6, FA, FB, MODE><<<grid, block, 0, s>>>(fc, a);
}

template <int FA, int FB, int MODE>
int launch_texpair_pipe_f(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win)
{
    vr::TexArgs a{};
    a.tex = c->tex2; a.out = d_out; a.local_rows = plan.local_rows;
    a.zlin = c->d_zlin; a.zpitch = (int)c->zpitch; a.zslice = (int)c->zslice;
    a.tf_lut = c->d_lut;
    const dim3 grid((c->W + 31) / 32, (plan.local_rows + 7) / 8);
    bool unit, recip, nocap;
    packed_flags(c, plan, &unit, &recip, &nocap);
    if (c->bpv == 2) {
        if (win == vr::WIN_COVERS0) launch_texpair_pipe_tw<uint16_t, vr::WIN_COVERS0, FA, FB, MODE>(plan.fc, a, grid, s, unit, recip, nocap);
        else                        launch_texpair_pipe_tw<uint16_t, vr::WIN_CLAMP, FA, FB, MODE>(plan.fc, a, grid, s, unit, recip, nocap);
    } else {
        if (win == vr::WIN_COVERS0) launch_texpair_pipe_tw<uint8_t, vr::WIN_COVERS0, FA, FB, MODE>(plan.fc, a, grid, s, unit, recip, nocap);
        else                        launch_texpair_pipe_tw<uint8_t, vr::WIN_CLAMP, FA, FB, MODE>(plan.fc, a, grid, s, unit, recip, nocap);
    }
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

// software-pipelined z-pair march (two fetches in flight per warp): texture gathers only, texture gather /
// LSU alternating (hybrid), or LSU only
int launch_texpair_pipe(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win, int kernel)
{
    if (kernel == VR_KERNEL_HYBRID) return launch_texpair_pipe_f<vr::FETCH_TEX, vr::FETCH_LSU, vr::MODE_DVR>(c, plan, d_out, s, win);
    if (kernel == VR_KERNEL_ZLSU)   return launch_texpair_pipe_f<vr::FETCH_LSU, vr::FETCH_LSU, vr::MODE_DVR>(c, plan, d_out, s, win);
    if (plan.fc.view_top)           return launch_texpair_pipe_f<vr::FETCH_TEX, vr::FETCH_TEX, vr::MODE_DVR_TOP>(c, plan, d_out, s, win);
    if (plan.fc.view_bottom)        return launch_texpair_pipe_f<vr::FETCH_TEX, vr::FETCH_TEX, vr::MODE_DVR_BOTTOM>(c, plan, d_out, s, win);
    if (plan.fc.is_mip)             return launch_texpair_pipe_f<vr::FETCH_TEX, vr::FETCH_TEX, vr::MODE_MIP>(c, plan, d_out, s, win);
    if (plan.fc.use_tf)             return launch_texpair_pipe_f<vr::FETCH_TEX, vr::FETCH_TEX, vr::MODE_TF>(c, plan, d_out, s, win);
    return launch_texpair_pipe_f<vr::FETCH_TEX, vr::FETCH_TEX, vr::MODE_DVR>(c, plan, d_out, s, win);
}

// two rays per thread: CTA = 64 x 8 pixels.  VR_TEXPAIR2_MINB=4 (lab) trades occupancy for registers.
int launch_texpair2(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win)
{
    vr::TexArgs a{};
    a.tex = c->tex2; a.out = d_out; a.local_rows = plan.local_rows;
    const dim3 grid((c->W + 63) / 64, (plan.local_rows + 7) / 8);
    bool unit, recip, nocap;
    packed_flags(c, plan, &unit, &recip, &nocap);
    static const int minb = [] { const char* e = std::getenv("VR_TEXPAIR2_MINB"); return e ? std::atoi(e) : 5; }();
#define VR_TP2(T, WIN) do { if (minb == 4) launch_texpair2_tw<T, WIN, 4>(plan.fc, a, grid, s, unit, recip, nocap); \
                            else           launch_texpair2_tw<T, WIN, 5>(plan.fc, a, grid, s, unit, recip, nocap); } while (0)
    if (c->bpv == 2) { if (win == vr::WIN_COVERS0) VR_TP2(uint16_t, vr::WIN_COVERS0); else VR_TP2(uint16_t, vr::WIN_CLAMP); }
    else             { if (win == vr::WIN_COVERS0) VR_TP2(uint8_t, vr::WIN_COVERS0);  else VR_TP2(uint8_t, vr::WIN_CLAMP); }
#undef VR_TP2
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

template <typename T>
int launch_fast_t(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, int win)
{
    using namespace vr;
    FastArgs a{};
    a.vol = c->d_vol; a.pitch = c->pitch; a.slice_lo = (uint32_t)c->slice; a.out = d_out; a.local_rows = plan.local_rows;
    const dim3 block(FAST_THREADS), grid((c->W + 31) / 32, (plan.local_rows + 7) / 8);
    const FrameConsts& fc = plan.fc;
    const bool tri = fc.filter == VR_FILTER_TRILINEAR, recip = plan.tcdiv == DIV_RECIP_EXACT, cov = win == WIN_COVERS0;
    const uint64_t padded_voxels = c->slice * (uint64_t)(c->dim[2] + 2);
    if (tri && padded_voxels < (1ull << 31)) {
        // trilinear: one ray per thread, f32x2 packing inside the ray (signed 32-bit texel indices)
        bool unit, recip2, nocap;
        packed_flags(c, plan, &unit, &recip2, &nocap);
        if (cov) launch_packed_tw<T, WIN_COVERS0>(fc, a, grid, s, unit, recip, nocap);
        else     launch_packed_tw<T, WIN_CLAMP>(fc, a, grid, s, unit, recip, nocap);
        VR_CUDA(cudaGetLastError());
        return VR_OK;
    }
#define VR_FAST(F, D, W) march_fast_kernel<T, F, D, W, FLOOR_XU1, 1><<<grid, block, 0, s>>>(fc, a)
    if (tri) {
        if (recip) { if (cov) VR_FAST(VR_FILTER_TRILINEAR, DIV_RECIP_EXACT, WIN_COVERS0); else VR_FAST(VR_FILTER_TRILINEAR, DIV_RECIP_EXACT, WIN_CLAMP); }
        else       { if (cov) VR_FAST(VR_FILTER_TRILINEAR, DIV_MARKSTEIN, WIN_COVERS0);   else VR_FAST(VR_FILTER_TRILINEAR, DIV_MARKSTEIN, WIN_CLAMP); }
    } else {
        if (recip) { if (cov) VR_FAST(VR_FILTER_NEAREST, DIV_RECIP_EXACT, WIN_COVERS0); else VR_FAST(VR_FILTER_NEAREST, DIV_RECIP_EXACT, WIN_CLAMP); }
        else       { if (cov) VR_FAST(VR_FILTER_NEAREST, DIV_MARKSTEIN, WIN_COVERS0);   else VR_FAST(VR_FILTER_NEAREST, DIV_MARKSTEIN, WIN_CLAMP); }
    }
#undef VR_FAST
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

// Linear copy of the z-pair words for the HYBRID / ZLSU lab kernels, built from the padded volume on
// first use: word (jx, jy, L) = padded(jx, jy, L) | padded(jx, jy, L+1) << bits, same row pitch.
template <typename T>
bool ensure_zlin_t(vr_context* c)
{
    typedef typename vr::PairWord<T>::type W;
    const uint64_t zslice = c->slice, zwords = zslice * (uint64_t)(c->dim[2] + 1);
    if (zwords + zslice >= (1ull << 31)) return false;              // 32-bit word indices in the kernel
    void* d_z = nullptr;
    if (cudaMalloc(&d_z, zwords * sizeof(W) + 256) != cudaSuccess) { cudaGetLastError(); return false; }
    vr::zpair_from_padded_kernel<T, W><<<c->sm_count * 16, 256, 0, c->stream>>>((const T*)c->d_vol, (W*)d_z, zslice, zwords);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { cudaFree(d_z); cudaGetLastError(); return false; }
    c->d_zlin = d_z; c->zpitch = c->pitch; c->zslice = zslice;
    return true;
}

bool ensure_zlin(vr_context* c)
{
    if (c->d_zlin) return true;
    if (!c->d_vol || !c->tex2) return false;
    return c->bpv == 1 ? ensure_zlin_t<uint8_t>(c) : ensure_zlin_t<uint16_t>(c);
}

// the march: returns which kernel ran
int launch_march(vr_context* c, const LaunchPlan& plan, float* d_out, cudaStream_t s, uint32_t* used,
                 uint32_t* launches)
{
    const vr::FrameConsts& fc = plan.fc;
    const uint64_t padded_voxels = c->slice * (uint64_t)(c->dim[2] + 2);
    // the optimised kernels cover: DVR, default view, no TF, ordered window with a verified
    // Markstein divisor, alpha_scale >= 0 (range tests on bit patterns), 32-bit texel indices,
    // correctly rounded tex-coord division without div.rn
    const bool base_ok = !plan.generic && plan.tcdiv != vr::DIV_IEEE && c->params.alpha_scale >= 0.0f;
    const bool fast_ok = base_ok && padded_voxels < (1ull << 32);          // LSU kernels: 32-bit texel indices
    const bool windowed_ok = fast_ok && vr::windowed_supported(fc, c->bpv, padded_voxels);
    const int win = (c->params.min_val == 0 && c->have_stats && c->stats.max_value <= c->params.max_val)
                        ? vr::WIN_COVERS0 : vr::WIN_CLAMP;
    const bool tex_ok = base_ok && fc.filter == VR_FILTER_TRILINEAR && c->tex != 0;   // hardware addressing: no index limit
    const bool texpair_ok = base_ok && fc.filter == VR_FILTER_TRILINEAR && c->tex2 != 0;
    const bool nearest_tex_ok = base_ok && fc.filter == VR_FILTER_NEAREST && c->tex != 0;
    int want = c->params.kernel;
    if (want == VR_KERNEL_NEAREST_TEX && !nearest_tex_ok) want = fast_ok ? VR_KERNEL_FAST : VR_KERNEL_DIRECT;
    if (want == VR_KERNEL_AUTO && nearest_tex_ok) want = VR_KERNEL_NEAREST_TEX;
    if ((fc.use_tf || fc.is_mip || fc.view_top || fc.view_bottom) && !plan.generic) want = texpair_ok ? VR_KERNEL_TEXPAIR_PIPE : VR_KERNEL_DIRECT;   // make_plan guarantees AUTO / TEXPAIR_PIPE here
    if (want == VR_KERNEL_AUTO) want = texpair_ok ? VR_KERNEL_TEXPAIR_PIPE : tex_ok ? VR_KERNEL_TEXGATHER : (fast_ok ? VR_KERNEL_FAST : VR_KERNEL_DIRECT);
    const bool zlin_ok = texpair_ok && (c->params.kernel == VR_KERNEL_HYBRID || c->params.kernel == VR_KERNEL_ZLSU) && ensure_zlin(c);
    if ((want == VR_KERNEL_HYBRID || want == VR_KERNEL_ZLSU) && !zlin_ok) want = VR_KERNEL_TEXPAIR_PIPE;
    if ((want == VR_KERNEL_TEXPAIR2 || want == VR_KERNEL_TEXPAIR_PIPE) && !texpair_ok) want = VR_KERNEL_TEXPAIR;
    if (want == VR_KERNEL_TEXPAIR && !texpair_ok) want = tex_ok ? VR_KERNEL_TEXGATHER : (fast_ok ? VR_KERNEL_FAST : VR_KERNEL_DIRECT);
    if (want == VR_KERNEL_TEXGATHER && !tex_ok) want = fast_ok ? VR_KERNEL_FAST : VR_KERNEL_DIRECT;
    if (want == VR_KERNEL_WINDOWED && !windowed_ok) want = fast_ok ? VR_KERNEL_FAST : VR_KERNEL_DIRECT;
    if (want == VR_KERNEL_FAST && !fast_ok) want = VR_KERNEL_DIRECT;
    *launches = 1;
    if (want == VR_KERNEL_WINDOWED) {
        int rc = c->bpv == 1
            ? vr::launch_windowed_t<uint8_t>(c->win, fc, c->d_vol, c->pitch, c->slice, c->dim[1], c->dim[2], d_out, plan.local_rows, c->sm_count, plan.tcdiv, win, s, false)
            : vr::launch_windowed_t<uint16_t>(c->win, fc, c->d_vol, c->pitch, c->slice, c->dim[1], c->dim[2], d_out, plan.local_rows, c->sm_count, plan.tcdiv, win, s, false);
        if (rc == 0) { *used = VR_KERNEL_WINDOWED; return VR_OK; }
        // no tensor map for this volume (e.g. smaller than one TMA box): use the L1 path
        cudaGetLastError();
        want = VR_KERNEL_FAST;
    }
    if (want == VR_KERNEL_NEAREST_TEX) {
        *used = VR_KERNEL_NEAREST_TEX;
        return launch_nearest_tex(c, plan, d_out, s, win);
    }
    if (want == VR_KERNEL_TEXPAIR_PIPE || want == VR_KERNEL_HYBRID || want == VR_KERNEL_ZLSU) {
        *used = (uint32_t)want;
        return launch_texpair_pipe(c, plan, d_out, s, win, want);
    }
    if (want == VR_KERNEL_TEXPAIR2) {
        *used = VR_KERNEL_TEXPAIR2;
        return launch_texpair2(c, plan, d_out, s, win);
    }
    if (want == VR_KERNEL_TEXPAIR) {
        *used = VR_KERNEL_TEXPAIR;
        return launch_texpair(c, plan, d_out, s, win);
    }
    if (want == VR_KERNEL_TEXGATHER) {
        *used = VR_KERNEL_TEXGATHER;
        return launch_texgather(c, plan, d_out, s, win);
    }
    if (want == VR_KERNEL_FAST) {
        *used = VR_KERNEL_FAST;
        return c->bpv == 1 ? launch_fast_t<uint8_t>(c, plan, d_out, s, win) : launch_fast_t<uint16_t>(c, plan, d_out, s, win);
    }
    *used = VR_KERNEL_DIRECT;
    return launch_direct(c, plan, d_out, s);
}

int render_common(vr_context* c, float* d_out, int compact, cudaStream_t s, vr_render_stats* stats)
{
    LaunchPlan plan;
    int rc = make_plan(c, compact, &plan);
    if (rc != VR_OK) return rc;
    uint32_t used = 0, launches = 0;
    VR_CUDA(cudaEventRecord(c->ev0, s));
    rc = launch_march(c, plan, d_out, s, &used, &launches);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaEventRecord(c->ev1, s));
    VR_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    VR_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (stats) { stats->kernel_ms = ms; stats->kernel_launches = launches; stats->kernel_used = used; }
    return VR_OK;
}

template <typename T>
int ingest_from_device(vr_context* c, const T* d_src, const uint64_t dims[3])
{
    const int nx = (int)dims[0], ny = (int)dims[1], nz = (int)dims[2];
    const uint64_t n = (uint64_t)nx * ny * nz;
    // stats first (RendererCore.cpp:360-405)
    unsigned int* d_mm = nullptr;
    unsigned long long* d_bins = nullptr;
    VR_CUDA(cudaMalloc(&d_mm, 2 * sizeof(unsigned int)));
    VR_CUDA(cudaMalloc(&d_bins, 256 * sizeof(unsigned long long)));
    const unsigned int init[2] = {0xffffffffu, 0u};
    VR_CUDA(cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    VR_CUDA(cudaMemsetAsync(d_bins, 0, 256 * sizeof(unsigned long long), c->stream));
    const int blocks = c->sm_count * 8;
    unsigned int mm[2] = {0, 255};
    if (sizeof(T) == 2) {
        vr::minmax_kernel<T><<<blocks, 256, 0, c->stream>>>(d_src, n, d_mm, d_mm + 1);
        VR_CUDA(cudaGetLastError());
        VR_CUDA(cudaMemcpyAsync(mm, d_mm, sizeof mm, cudaMemcpyDeviceToHost, c->stream));
        VR_CUDA(cudaStreamSynchronize(c->stream));
    }
    {
        uint8_t* d_hlut = nullptr;
        size_t smem = (vr::HIST_THREADS / 32) * 256 * sizeof(unsigned int);
        if (sizeof(T) == 2) {
            VR_CUDA(cudaMalloc(&d_hlut, 65536));
            vr::histogram_lut_kernel<<<65536 / 256, 256, 0, c->stream>>>((float)(int)mm[1], d_hlut);
            smem += 65536;
            VR_CUDA(cudaFuncSetAttribute(vr::histogram_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        vr::histogram_kernel<T><<<c->sm_count * 3, vr::HIST_THREADS, smem, c->stream>>>(d_src, n, d_hlut, d_bins);
        cudaError_t he = cudaGetLastError();
        if (he == cudaSuccess) he = cudaStreamSynchronize(c->stream);
        if (d_hlut) cudaFree(d_hlut);
        if (he != cudaSuccess) { cudaFree(d_mm); cudaFree(d_bins); return cuda_fail(he, "histogram_kernel"); }
    }
    unsigned long long bins[256];
    VR_CUDA(cudaMemcpyAsync(bins, d_bins, sizeof bins, cudaMemcpyDeviceToHost, c->stream));

    // padded copy
    const uint32_t pitch = (uint32_t)(round_up((uint64_t)(nx + 2) * sizeof(T), 16) / sizeof(T));
    const uint64_t slice = (uint64_t)pitch * (uint64_t)(ny + 2);
    const uint64_t bytes = slice * (uint64_t)(nz + 2) * sizeof(T) + 256;
    void* d_new = nullptr;
    cudaError_t e = cudaMalloc(&d_new, bytes);
    if (e != cudaSuccess) { cudaFree(d_mm); cudaFree(d_bins); return cuda_fail(e, "cudaMalloc(padded volume)"); }
    VR_CUDA(cudaMemsetAsync(d_new, 0, bytes, c->stream));
    vr::pad_volume_kernel<T><<<c->sm_count * 16, 256, 0, c->stream>>>(d_src, (T*)d_new, nx, ny, nz, pitch);
    VR_CUDA(cudaGetLastError());
    VR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_mm); cudaFree(d_bins);

    // layered array for the gather kernel (limits of layered 2-D arrays: 32768 x 32768 x 2048)
    if (c->tex) { cudaDestroyTextureObject(c->tex); c->tex = 0; }
    if (c->d_arr) { cudaFreeArray(c->d_arr); c->d_arr = nullptr; }
    if (nz <= 2048 && nx <= 32768 && ny <= 32768) {
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(8 * (int)sizeof(T), 0, 0, 0, cudaChannelFormatKindUnsigned);
        cudaArray_t arr = nullptr;
        if (cudaMalloc3DArray(&arr, &cd, make_cudaExtent(nx, ny, nz), cudaArrayLayered) == cudaSuccess) {
            cudaMemcpy3DParms cp = {};
            cp.srcPtr = make_cudaPitchedPtr(const_cast<T*>(d_src), (size_t)nx * sizeof(T), nx, ny);
            cp.dstArray = arr; cp.extent = make_cudaExtent(nx, ny, nz); cp.kind = cudaMemcpyDeviceToDevice;
            cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
            cudaTextureObject_t tex = 0;
            if (cudaMemcpy3DAsync(&cp, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess &&
                cudaCreateTextureObject(&tex, &rd, &td, nullptr) == cudaSuccess) {
                c->d_arr = arr; c->tex = tex;
            } else {
                cudaFreeArray(arr);
            }
        }
        cudaGetLastError();     // without the array the LDG kernels serve every frame
    }
    // z-pair array for the texpair kernel: Nz+1 layers of (z, z+1) words, packed and copied in
    // chunks of layers through a bounded staging buffer
    if (c->tex2) { cudaDestroyTextureObject(c->tex2); c->tex2 = 0; }
    if (c->d_arr2) { cudaFreeArray(c->d_arr2); c->d_arr2 = nullptr; }
    if (nz + 1 <= 2048 && nx <= 32768 && ny <= 32768) {
        typedef typename vr::PairWord<T>::type W;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(8 * (int)sizeof(W), 0, 0, 0, cudaChannelFormatKindUnsigned);
        cudaArray_t arr = nullptr;
        W* d_stage = nullptr;
        const uint64_t per_layer = (uint64_t)nx * ny;
        const int chunk = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)nz + 1, (256ull << 20) / (per_layer * sizeof(W))));
        bool ok = cudaMalloc3DArray(&arr, &cd, make_cudaExtent(nx, ny, nz + 1), cudaArrayLayered) == cudaSuccess &&
                  cudaMalloc(&d_stage, per_layer * (uint64_t)chunk * sizeof(W)) == cudaSuccess;
        for (int l0 = 0; ok && l0 <= nz; l0 += chunk) {
            const int nl = std::min(chunk, nz + 1 - l0);
            vr::zpair_pack_kernel<T, W><<<c->sm_count * 16, 256, 0, c->stream>>>(d_src, d_stage, nx, ny, nz, l0, nl);
            cudaMemcpy3DParms cp = {};
            cp.srcPtr = make_cudaPitchedPtr(d_stage, (size_t)nx * sizeof(W), nx, ny);
            cp.dstArray = arr; cp.dstPos = make_cudaPos(0, 0, l0);
            cp.extent = make_cudaExtent(nx, ny, nl); cp.kind = cudaMemcpyDeviceToDevice;
            ok = cudaGetLastError() == cudaSuccess && cudaMemcpy3DAsync(&cp, c->stream) == cudaSuccess;
        }
        ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
        if (d_stage) cudaFree(d_stage);
        cudaTextureObject_t tex = 0;
        if (ok) {
            cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
            ok = cudaCreateTextureObject(&tex, &rd, &td, nullptr) == cudaSuccess;
        }
        if (ok) { c->d_arr2 = arr; c->tex2 = tex; }
        else if (arr) cudaFreeArray(arr);
        cudaGetLastError();     // without the array the other kernels serve every frame
    }
    if (c->d_zlin) { cudaFree(c->d_zlin); c->d_zlin = nullptr; }      // rebuilt on demand (ensure_zlin)
    if (c->d_vol) cudaFree(c->d_vol);
    c->d_vol = d_new; c->vol_bytes = bytes;
    c->dim[0] = nx; c->dim[1] = ny; c->dim[2] = nz;
    c->bpv = (int)sizeof(T); c->pitch = pitch; c->slice = slice;
    vr::windowed_invalidate(c->win);

    // histogram normalisation exactly as RendererCore.cpp:361,386-405: float bins that are
    // incremented one by one saturate at 2^24; max_value starts at the 16-bit dataset max
    // (or -1 for 8-bit data) and is then raised by the bin counts.
    vr_volume_stats& st = c->stats;
    if (sizeof(T) == 2) { st.min_value = (int)mm[0]; st.max_value = (int)mm[1]; }
    else { st.min_value = 0; st.max_value = 255; }
    int max_value = sizeof(T) == 2 ? (int)mm[1] : -1;
    float hist[256];
    for (int i = 0; i < 256; ++i) {
        const unsigned long long cnt = bins[i] > 16777216ull ? 16777216ull : bins[i];
        hist[i] = (float)cnt;
        if (i > 0 && hist[i] > (float)max_value) max_value = (int)hist[i];
    }
    for (int i = 0; i < 256; ++i) st.histogram[i] = hist[i] * 100.0f / (float)max_value;
    c->have_stats = true;
    return VR_OK;
}

int check_dims(const uint64_t dims[3], int bpv, const float voxel_size[3])
{
    if (!dims || !voxel_size) return fail(VR_ERR_INVALID, "upload: null argument");
    if (bpv != 1 && bpv != 2) return fail(VR_ERR_INVALID, "upload: bytes_per_voxel must be 1 or 2");
    for (int i = 0; i < 3; ++i) {
        if (dims[i] < 1 || dims[i] > 16384) return fail(VR_ERR_INVALID, "upload: each dimension must be in [1,16384]");
        if (!(voxel_size[i] > 0.0f) || !std::isfinite(voxel_size[i])) return fail(VR_ERR_INVALID, "upload: voxel_size must be finite and > 0");
    }
    return VR_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------ ABI

extern "C" {

const char* vr_version(void) { return "volren_b200 0.1 (sm_100a)"; }
const char* vr_last_error(void) { return g_last_error.c_str(); }

void vr_params_default(vr_params* p)
{
    if (!p) return;
    std::memset(p, 0, sizeof *p);
    p->alpha_scale = 1.0f;          // RendererCore.cpp:18
    p->min_val = 0; p->max_val = 0; // RendererCore.cpp:19-20
    p->filter = VR_FILTER_NEAREST;
    p->step_scale = 1.0f;
    p->kernel = VR_KERNEL_AUTO;
}

int vr_device_count(int* count)
{
    if (!count) return fail(VR_ERR_INVALID, "vr_device_count: null");
    VR_CUDA(cudaGetDeviceCount(count));
    return VR_OK;
}

int vr_create(int device, int width, int height, vr_context** out)
{
    if (!out) return fail(VR_ERR_INVALID, "vr_create: null out");
    *out = nullptr;
    if (width < 1 || height < 1 || width > 32768 || height > 32768)
        return fail(VR_ERR_INVALID, "vr_create: image size out of range");
    int n = 0;
    VR_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(VR_ERR_INVALID, "vr_create: no such CUDA device");
    VR_CUDA(cudaSetDevice(device));
    vr_context* c = new (std::nothrow) vr_context();
    if (!c) return fail(VR_ERR_OOM, "vr_create: out of host memory");
    c->device = device; c->W = width; c->H = height;
    vr_params_default(&c->params);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_frame, frame_alloc_bytes(width, height));
    if (e == cudaSuccess) e = cudaMemset(c->d_frame + (size_t)width * height * 4, 0, FRAME_SYNC_BYTES);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_lut, 256 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_flag, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->d_lut, 0, 256 * sizeof(float));
    if (e != cudaSuccess) { int rc = cuda_fail(e, "vr_create"); vr_destroy(c); return rc; }
    *out = c;
    return VR_OK;
}

void vr_destroy(vr_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    vr::windowed_release(c->win);
    if (c->tex) cudaDestroyTextureObject(c->tex);
    if (c->d_arr) cudaFreeArray(c->d_arr);
    if (c->tex2) cudaDestroyTextureObject(c->tex2);
    if (c->d_arr2) cudaFreeArray(c->d_arr2);
    if (c->d_zlin) cudaFree(c->d_zlin);
    if (c->d_vol) cudaFree(c->d_vol);
    if (c->d_frame) cudaFree(c->d_frame);
    if (c->d_rgb8) cudaFree(c->d_rgb8);
    if (c->d_lut) cudaFree(c->d_lut);
    if (c->d_flag) cudaFree(c->d_flag);
    for (int b = 0; b < vr_context::BANDS; ++b) {
        if (c->band_stream[b]) cudaStreamDestroy(c->band_stream[b]);
        if (c->band_kdone[b]) cudaEventDestroy(c->band_kdone[b]);
        if (c->band_cdone[b]) cudaEventDestroy(c->band_cdone[b]);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int vr_resize(vr_context* c, int width, int height)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_resize: null context");
    if (width < 1 || height < 1 || width > 32768 || height > 32768)
        return fail(VR_ERR_INVALID, "vr_resize: image size out of range");
    VR_CUDA(cudaSetDevice(c->device));
    float* d_new = nullptr;
    VR_CUDA(cudaMalloc(&d_new, frame_alloc_bytes(width, height)));
    VR_CUDA(cudaMemset(d_new + (size_t)width * height * 4, 0, FRAME_SYNC_BYTES));
    cudaFree(c->d_frame);
    if (c->d_rgb8) { cudaFree(c->d_rgb8); c->d_rgb8 = nullptr; }
    c->d_frame = d_new; c->W = width; c->H = height;
    return VR_OK;
}

int vr_image_size(const vr_context* c, int* width, int* height)
{
    if (!c || !width || !height) return fail(VR_ERR_INVALID, "vr_image_size: null");
    *width = c->W; *height = c->H;
    return VR_OK;
}

int vr_upload_volume(vr_context* c, const void* voxels, const uint64_t dims[3], int bpv, const float voxel_size[3])
{
    if (!c || !voxels) return fail(VR_ERR_INVALID, "vr_upload_volume: null argument");
    int rc = check_dims(dims, bpv, voxel_size);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(c->device));
    const uint64_t bytes = dims[0] * dims[1] * dims[2] * (uint64_t)bpv;
    void* d_src = nullptr;
    VR_CUDA(cudaMalloc(&d_src, bytes));
    cudaError_t e = cudaMemcpyAsync(d_src, voxels, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { cudaFree(d_src); return cuda_fail(e, "vr_upload_volume: H2D copy"); }
    rc = bpv == 1 ? ingest_from_device<uint8_t>(c, (const uint8_t*)d_src, dims)
                  : ingest_from_device<uint16_t>(c, (const uint16_t*)d_src, dims);
    cudaFree(d_src);
    if (rc != VR_OK) return rc;
    for (int i = 0; i < 3; ++i) c->voxel_size[i] = voxel_size[i];
    return VR_OK;
}

int vr_upload_volume_device(vr_context* c, const void* d_voxels, const uint64_t dims[3], int bpv, const float voxel_size[3])
{
    if (!c || !d_voxels) return fail(VR_ERR_INVALID, "vr_upload_volume_device: null argument");
    int rc = check_dims(dims, bpv, voxel_size);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(c->device));
    VR_CUDA(cudaDeviceSynchronize());   // the source may have been produced on another stream
    rc = bpv == 1 ? ingest_from_device<uint8_t>(c, (const uint8_t*)d_voxels, dims)
                  : ingest_from_device<uint16_t>(c, (const uint16_t*)d_voxels, dims);
    if (rc != VR_OK) return rc;
    for (int i = 0; i < 3; ++i) c->voxel_size[i] = voxel_size[i];
    return VR_OK;
}

int vr_set_voxel_size(vr_context* c, const float voxel_size[3])
{
    if (!c || !voxel_size) return fail(VR_ERR_INVALID, "vr_set_voxel_size: null");
    for (int i = 0; i < 3; ++i)
        if (!(voxel_size[i] > 0.0f) || !std::isfinite(voxel_size[i]))
            return fail(VR_ERR_INVALID, "vr_set_voxel_size: must be finite and > 0");
    for (int i = 0; i < 3; ++i) c->voxel_size[i] = voxel_size[i];
    return VR_OK;
}

int vr_volume_stats_get(vr_context* c, vr_volume_stats* out)
{
    if (!c || !out) return fail(VR_ERR_INVALID, "vr_volume_stats_get: null");
    if (!c->have_stats) return fail(VR_ERR_NO_VOLUME, "vr_volume_stats_get: no volume uploaded");
    *out = c->stats;
    return VR_OK;
}

int vr_set_camera(vr_context* c, const float cam21[21])
{
    if (!c || !cam21) return fail(VR_ERR_INVALID, "vr_set_camera: null");
    for (int i = 0; i < 21; ++i)
        if (!std::isfinite(cam21[i])) return fail(VR_ERR_INVALID, "vr_set_camera: non-finite camera block");
    std::memcpy(c->cam, cam21, sizeof(float) * 21);
    c->have_cam = true;
    return VR_OK;
}

int vr_set_params(vr_context* c, const vr_params* p)
{
    if (!c || !p) return fail(VR_ERR_INVALID, "vr_set_params: null");
    if (p->filter != VR_FILTER_NEAREST && p->filter != VR_FILTER_TRILINEAR)
        return fail(VR_ERR_INVALID, "vr_set_params: unknown filter");
    if (!std::isfinite(p->alpha_scale)) return fail(VR_ERR_INVALID, "vr_set_params: alpha_scale not finite");
    if (!(p->step_scale > 0.0f) || !std::isfinite(p->step_scale))
        return fail(VR_ERR_INVALID, "vr_set_params: step_scale must be finite and > 0");
    if (p->kernel < VR_KERNEL_AUTO || p->kernel > VR_KERNEL_NEAREST_TEX)
        return fail(VR_ERR_INVALID, "vr_set_params: unknown kernel");
    if (p->use_tf) {
        // the optimised loop tests ranges on float bit patterns: it needs a finite, non-negative opacity LUT
        bool ok = true;
        for (int i = 0; i < 256; ++i) ok = ok && std::isfinite(p->tf_lut[i]) && p->tf_lut[i] >= 0.0f && !std::signbit(p->tf_lut[i]);
        c->lut_fast_ok = ok;
        VR_CUDA(cudaSetDevice(c->device));
        VR_CUDA(cudaMemcpyAsync(c->d_lut, p->tf_lut, 256 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        VR_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->params = *p;
    return VR_OK;
}

int vr_get_params(const vr_context* c, vr_params* p)
{
    if (!c || !p) return fail(VR_ERR_INVALID, "vr_get_params: null");
    *p = c->params;
    return VR_OK;
}

int vr_set_partition(vr_context* c, int rank, int world, int tile_rows)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_set_partition: null context");
    if (world < 1 || rank < 0 || rank >= world || tile_rows < 1)
        return fail(VR_ERR_INVALID, "vr_set_partition: need 0 <= rank < world and tile_rows >= 1");
    c->rank = rank; c->world = world; c->tile_rows = tile_rows;
    return VR_OK;
}

int vr_owned_rows(const vr_context* c, int* rows)
{
    if (!c || !rows) return fail(VR_ERR_INVALID, "vr_owned_rows: null");
    *rows = max_compact_rows(c->H, c->world, c->tile_rows);
    return VR_OK;
}

int vr_render_device(vr_context* c, float* d_rgba, int compact, void* cuda_stream, vr_render_stats* stats)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_render_device: null context");
    if (!d_rgba) { d_rgba = c->d_frame; compact = 0; }
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    int rc = render_common(c, d_rgba, compact ? 1 : 0, s, stats);
    if (rc != VR_OK) return rc;
    if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

// vr_render for an unpartitioned frame: BANDS horizontal bands, each marched on its own stream and
// followed there by the device->host copy of its rows, so only the last band's copy is exposed
// (33 MB over PCIe cost 0.7 ms per 1080p frame when copied after the whole march).  A band is rendered
// through the row-tile partition (rank = band, world = BANDS, one tile per band): no kernel changes.
static int render_banded(vr_context* c, float* host_rgba, vr_render_stats* stats)
{
    constexpr int B = vr_context::BANDS;
    if (!c->bands_ready) {
        for (int b = 0; b < B; ++b) {
            VR_CUDA(cudaStreamCreateWithFlags(&c->band_stream[b], cudaStreamNonBlocking));
            VR_CUDA(cudaEventCreateWithFlags(&c->band_kdone[b], cudaEventDisableTiming));
            VR_CUDA(cudaEventCreateWithFlags(&c->band_cdone[b], cudaEventDisableTiming));
        }
        c->bands_ready = true;
    }
    const int band_rows = (int)round_up((uint64_t)(c->H + B - 1) / B, 8);
    const int save_rank = c->rank, save_world = c->world, save_tile = c->tile_rows;
    uint32_t used = 0, launches = 0, total_launches = 0;
    int rc = VR_OK, nb = 0;
    cudaError_t e = cudaEventRecord(c->ev0, c->stream);
    for (int b = 0; b < B && rc == VR_OK && e == cudaSuccess; ++b) {
        const int y0 = b * band_rows, y1 = std::min(c->H, y0 + band_rows);
        if (y0 >= c->H) break;
        c->rank = b; c->world = B; c->tile_rows = band_rows;
        LaunchPlan plan;
        rc = make_plan(c, 0, &plan);
        if (rc != VR_OK) break;
        cudaStream_t bs = c->band_stream[b];
        e = cudaStreamWaitEvent(bs, c->ev0, 0);
        if (e == cudaSuccess) rc = launch_march(c, plan, c->d_frame, bs, &used, &launches);
        total_launches += launches;
        if (e == cudaSuccess) e = cudaEventRecord(c->band_kdone[b], bs);
        const size_t off = (size_t)y0 * c->W * 4, n = (size_t)(y1 - y0) * c->W * 4 * sizeof(float);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host_rgba + off, c->d_frame + off, n, cudaMemcpyDeviceToHost, bs);
        if (e == cudaSuccess) e = cudaEventRecord(c->band_cdone[b], bs);
        nb = b + 1;
    }
    c->rank = save_rank; c->world = save_world; c->tile_rows = save_tile;
    for (int b = 0; b < nb && e == cudaSuccess; ++b) e = cudaStreamWaitEvent(c->stream, c->band_kdone[b], 0);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev1, c->stream);           // every band's march has finished
    for (int b = 0; b < nb && e == cudaSuccess; ++b) e = cudaStreamWaitEvent(c->stream, c->band_cdone[b], 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        for (int b = 0; b < nb; ++b) cudaStreamSynchronize(c->band_stream[b]);
        return cuda_fail(e, "vr_render (banded)");
    }
    if (rc != VR_OK) { for (int b = 0; b < nb; ++b) cudaStreamSynchronize(c->band_stream[b]); return rc; }
    float ms = 0.f;
    VR_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (stats) { stats->kernel_ms = ms; stats->kernel_launches = total_launches; stats->kernel_used = used; }
    return VR_OK;
}

int vr_render(vr_context* c, float* host_rgba, vr_render_stats* stats)
{
    if (!c || !host_rgba) return fail(VR_ERR_INVALID, "vr_render: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    const size_t bytes = (size_t)c->W * c->H * 4 * sizeof(float);
    // stateless kernels only (the windowed kernel keeps per-context scratch)
    static const bool no_bands = std::getenv("VR_NO_BANDS") != nullptr;
    // small frames copy in microseconds: four launches on four streams would cost more than they hide
    const bool worth_banding = (size_t)c->W * c->H >= ((size_t)1 << 19);
    if (!no_bands && worth_banding && c->world == 1 && c->H >= 8 * vr_context::BANDS && c->params.kernel != VR_KERNEL_WINDOWED) {
        int rc = render_banded(c, host_rgba, stats);
        if (rc != VR_OK) return rc;
        if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return VR_OK;
    }
    if (c->world > 1) VR_CUDA(cudaMemsetAsync(c->d_frame, 0, bytes, c->stream));
    int rc = render_common(c, c->d_frame, 0, c->stream, stats);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaMemcpyAsync(host_rgba, c->d_frame, bytes, cudaMemcpyDeviceToHost, c->stream));
    VR_CUDA(cudaStreamSynchronize(c->stream));
    if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

// Multi-GPU end to end: this rank's row tiles straight into a FULL host frame (one buffer shared by all
// ranks, e.g. POSIX shared memory registered with cudaHostRegister in every process): N PCIe links carry
// the frame instead of rank 0's one, and nothing crosses NVLink.  Rows this rank does not own are not touched.
int vr_render_owned_to_host(vr_context* c, float* host_full_frame, vr_render_stats* stats)
{
    if (!c || !host_full_frame) return fail(VR_ERR_INVALID, "vr_render_owned_to_host: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    int rc = render_common(c, c->d_frame, /*compact=*/1, c->stream, stats);     // compact tiles fit in the frame buffer
    if (rc != VR_OK) return rc;
    const int tiles = (c->H + c->tile_rows - 1) / c->tile_rows;
    const size_t row_floats = (size_t)c->W * 4;
    int local_tile = 0;
    for (int t = c->rank; t < tiles; t += c->world, ++local_tile) {
        const int y0 = t * c->tile_rows, rows = std::min(c->tile_rows, c->H - y0);
        VR_CUDA(cudaMemcpyAsync(host_full_frame + (size_t)y0 * row_floats,
                                c->d_frame + (size_t)local_tile * c->tile_rows * row_floats,
                                (size_t)rows * row_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    VR_CUDA(cudaStreamSynchronize(c->stream));
    if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

int vr_read_frame(vr_context* c, float* host_rgba)
{
    if (!c || !host_rgba) return fail(VR_ERR_INVALID, "vr_read_frame: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    VR_CUDA(cudaMemcpyAsync(host_rgba, c->d_frame, (size_t)c->W * c->H * 4 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    VR_CUDA(cudaStreamSynchronize(c->stream));
    return VR_OK;
}

int vr_assemble_tiles(vr_context* c, const float* d_gathered, float* d_frame, int world, int tile_rows, void* cuda_stream)
{
    if (!c || !d_gathered || !d_frame || world < 1 || tile_rows < 1)
        return fail(VR_ERR_INVALID, "vr_assemble_tiles: bad argument");
    VR_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    const int rows_per_rank = max_compact_rows(c->H, world, tile_rows);
    const dim3 block(128), grid((c->W + 127) / 128, c->H);
    vr::assemble_tiles_kernel<<<grid, block, 0, s>>>((const float4*)d_gathered, (float4*)d_frame,
                                                     c->W, c->H, world, tile_rows, rows_per_rank);
    VR_CUDA(cudaGetLastError());
    VR_CUDA(cudaStreamSynchronize(s));
    return VR_OK;
}

int vr_frame_device_ptr(vr_context* c, float** d_frame)
{
    if (!c || !d_frame) return fail(VR_ERR_INVALID, "vr_frame_device_ptr: null argument");
    *d_frame = c->d_frame;
    return VR_OK;
}

int vr_frame_export_ipc(vr_context* c, unsigned char handle[64])
{
    if (!c || !handle) return fail(VR_ERR_INVALID, "vr_frame_export_ipc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    VR_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    VR_CUDA(cudaIpcGetMemHandle(&h, c->d_frame));
    std::memcpy(handle, &h, 64);
    return VR_OK;
}

int vr_frame_open_ipc(vr_context* c, const unsigned char handle[64], float** d_peer_frame)
{
    if (!c || !handle || !d_peer_frame) return fail(VR_ERR_INVALID, "vr_frame_open_ipc: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* p = nullptr;
    VR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_peer_frame = static_cast<float*>(p);
    return VR_OK;
}

int vr_frame_close_ipc(vr_context* c, float* d_peer_frame)
{
    if (!c || !d_peer_frame) return fail(VR_ERR_INVALID, "vr_frame_close_ipc: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    VR_CUDA(cudaIpcCloseMemHandle(d_peer_frame));
    return VR_OK;
}

// bound of the barrier spins: 5 s by default (a frame takes milliseconds); VR_PEER_TIMEOUT_MS overrides it,
// e.g. for rank processes that time-share ONE GPU, where a spinning wait kernel holds its whole time slice
static unsigned long long peer_timeout_ns()
{
    static const unsigned long long ns = [] {
        const char* e = std::getenv("VR_PEER_TIMEOUT_MS");
        const long long ms = e ? std::atoll(e) : 5000;
        return (unsigned long long)(ms > 0 ? ms : 5000) * 1000000ull;
    }();
    return ns;
}

// the barrier words live behind the pixels of the frame `d_target_frame` points to (same W x H on every rank)
static unsigned int* frame_sync_words(const vr_context* c, float* d_target_frame)
{
    return reinterpret_cast<unsigned int*>(d_target_frame + (size_t)c->W * c->H * 4);
}

int vr_peer_frame_arrive(vr_context* c, float* d_target_frame, uint32_t frame_no, int world, int is_owner, void* cuda_stream)
{
    if (!c || !d_target_frame || world < 1) return fail(VR_ERR_INVALID, "vr_peer_frame_arrive: bad argument");
    VR_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream;
    unsigned int* sync = frame_sync_words(c, d_target_frame);
    vr::peer_signal_kernel<<<1, 1, 0, s>>>(sync);
    if (is_owner) vr::peer_wait_kernel<<<1, 1, 0, s>>>(sync, 0, frame_no * (uint32_t)world, peer_timeout_ns());
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

int vr_peer_frame_release(vr_context* c, float* d_target_frame, uint32_t frame_no, int is_owner, void* cuda_stream)
{
    if (!c || !d_target_frame) return fail(VR_ERR_INVALID, "vr_peer_frame_release: bad argument");
    VR_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream;
    unsigned int* sync = frame_sync_words(c, d_target_frame);
    if (is_owner) vr::peer_release_kernel<<<1, 1, 0, s>>>(sync, frame_no);
    else          vr::peer_wait_kernel<<<1, 1, 0, s>>>(sync, 1, frame_no, peer_timeout_ns());
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

int vr_peer_frame_status(vr_context* c, float* d_target_frame, uint32_t* arrivals, uint32_t* released, uint32_t* timed_out)
{
    if (!c || !d_target_frame) return fail(VR_ERR_INVALID, "vr_peer_frame_status: bad argument");
    VR_CUDA(cudaSetDevice(c->device));
    unsigned int w[3] = {0, 0, 0};
    VR_CUDA(cudaMemcpy(w, frame_sync_words(c, d_target_frame), sizeof w, cudaMemcpyDeviceToHost));
    if (arrivals) *arrivals = w[0];
    if (released) *released = w[1];
    if (timed_out) *timed_out = w[2];
    return VR_OK;
}

int vr_read_rgb8(vr_context* c, uint8_t* host_rgb, int flip_vertical)
{
    if (!c || !host_rgb) return fail(VR_ERR_INVALID, "vr_read_rgb8: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)c->W * c->H * 3;
    if (!c->d_rgb8) VR_CUDA(cudaMalloc(&c->d_rgb8, bytes));
    const dim3 block(128), grid((c->W + 127) / 128, c->H);
    vr::rgba32f_to_rgb8_kernel<<<grid, block, 0, c->stream>>>((const float4*)c->d_frame, c->d_rgb8, c->W, c->H, flip_vertical ? 1 : 0);
    VR_CUDA(cudaGetLastError());
    VR_CUDA(cudaMemcpyAsync(host_rgb, c->d_rgb8, bytes, cudaMemcpyDeviceToHost, c->stream));
    VR_CUDA(cudaStreamSynchronize(c->stream));
    return VR_OK;
}

int vr_count_frame(vr_context* c, uint64_t* distinct_voxels, uint64_t* samples, uint64_t* rays_hit)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_count_frame: null context");
    VR_CUDA(cudaSetDevice(c->device));
    LaunchPlan plan;
    int rc = make_plan(c, 0, &plan);
    if (rc != VR_OK) return rc;
    const uint64_t nvox = (uint64_t)c->dim[0] * c->dim[1] * c->dim[2];
    const uint64_t nwords = (nvox + 31) / 32;
    unsigned int* d_bits = nullptr;
    unsigned long long* d_cnt = nullptr;
    VR_CUDA(cudaMalloc(&d_bits, nwords * sizeof(unsigned int)));
    cudaError_t e = cudaMalloc(&d_cnt, 3 * sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(d_bits); return cuda_fail(e, "vr_count_frame: cudaMalloc"); }
    cudaMemsetAsync(d_bits, 0, nwords * sizeof(unsigned int), c->stream);
    cudaMemsetAsync(d_cnt, 0, 3 * sizeof(unsigned long long), c->stream);
    vr::DirectArgs args{};
    args.vol = c->d_vol; args.pitch = c->pitch; args.slice = c->slice;
    args.tf_lut = c->d_lut; args.out = c->d_frame; args.local_rows = plan.local_rows;
    args.touch_bits = d_bits; args.counters = d_cnt;
    rc = c->bpv == 1 ? launch_direct_t<uint8_t, true>(c, plan, args, c->stream)
                     : launch_direct_t<uint16_t, true>(c, plan, args, c->stream);
    if (rc == VR_OK) {
        vr::popcount_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_bits, nwords, d_cnt + 2);
        unsigned long long h[3] = {0, 0, 0};
        e = cudaMemcpyAsync(h, d_cnt, sizeof h, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "vr_count_frame");
        if (samples) *samples = h[0];
        if (rays_hit) *rays_hit = h[1];
        if (distinct_voxels) *distinct_voxels = h[2];
    }
    cudaFree(d_bits); cudaFree(d_cnt);
    return rc;
}

static int synth_into(int sm_count, cudaStream_t s, void* d_dst, const uint64_t dims[3], int bpv,
                      uint32_t vmax, uint32_t seed, int with_hash)
{
    const int blocks = sm_count * 16;
    if (bpv == 1)
        vr::synth_mix_kernel<uint8_t><<<blocks, 256, 0, s>>>((uint8_t*)d_dst, (int)dims[0], (int)dims[1], (int)dims[2], vmax, seed, with_hash);
    else
        vr::synth_mix_kernel<uint16_t><<<blocks, 256, 0, s>>>((uint16_t*)d_dst, (int)dims[0], (int)dims[1], (int)dims[2], vmax, seed, with_hash);
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

int vr_upload_synthetic(vr_context* c, const uint64_t dims[3], int bpv, const float voxel_size[3],
                        uint32_t vmax, uint32_t seed, int with_hash_noise, void* d_copy_out)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_upload_synthetic: null context");
    int rc = check_dims(dims, bpv, voxel_size);
    if (rc != VR_OK) return rc;
    if (vmax > (bpv == 1 ? 255u : 65535u)) return fail(VR_ERR_INVALID, "vr_upload_synthetic: vmax too large");
    VR_CUDA(cudaSetDevice(c->device));
    const uint64_t bytes = dims[0] * dims[1] * dims[2] * (uint64_t)bpv;
    void* d_src = nullptr;
    VR_CUDA(cudaMalloc(&d_src, bytes));
    rc = synth_into(c->sm_count, c->stream, d_src, dims, bpv, vmax, seed, with_hash_noise);
    if (rc == VR_OK)
        rc = bpv == 1 ? ingest_from_device<uint8_t>(c, (const uint8_t*)d_src, dims)
                      : ingest_from_device<uint16_t>(c, (const uint16_t*)d_src, dims);
    if (rc == VR_OK && d_copy_out) {
        cudaError_t e = cudaMemcpy(d_copy_out, d_src, bytes, cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) rc = cuda_fail(e, "vr_upload_synthetic: copy out");
    }
    cudaFree(d_src);
    if (rc != VR_OK) return rc;
    for (int i = 0; i < 3; ++i) c->voxel_size[i] = voxel_size[i];
    return VR_OK;
}

int vr_synthetic_to_host(int device, const uint64_t dims[3], int bpv, uint32_t vmax, uint32_t seed,
                         int with_hash_noise, void* host_out)
{
    if (!host_out) return fail(VR_ERR_INVALID, "vr_synthetic_to_host: null");
    const float one[3] = {1.f, 1.f, 1.f};
    int rc = check_dims(dims, bpv, one);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VR_CUDA(cudaGetDeviceProperties(&prop, device));
    const uint64_t bytes = dims[0] * dims[1] * dims[2] * (uint64_t)bpv;
    void* d = nullptr;
    VR_CUDA(cudaMalloc(&d, bytes));
    rc = synth_into(prop.multiProcessorCount, nullptr, d, dims, bpv, vmax, seed, with_hash_noise);
    if (rc == VR_OK) {
        cudaError_t e = cudaMemcpy(host_out, d, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "vr_synthetic_to_host: D2H");
    }
    cudaFree(d);
    return rc;
}

}  // extern "C"
