// volren_abi.cu -- context management and the C-ABI of libvolren_b200.so (include/volren_b200.h).
//
// Host code in this file evaluates the per-frame constants (frame.h) and must be compiled
// with -Xcompiler -ffp-contract=off.  There is no CPU fallback anywhere in this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
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
#include "kernel_march.cuh"
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

// device allocation released on scope exit unless handed over (upload error paths must not leak)
struct DevBuf {
    void* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    void* release() { void* q = p; p = nullptr; return q; }
    template <typename U> U* as() const { return static_cast<U*>(p); }
};

// pipeline depth / resident CTAs per SM of the optimised kernels, measured on the headline frame
// (lab r2: depth 4 at 64 registers / 4 CTAs per SM is the best choice across cameras, partitions, volume
// shapes; the skipping forms carry the leap state and want the same register budget)
// CTA shape (lab r2): 4 warps covering 16 x 8 pixels.  Against the 8-warp 32 x 8 shape: -2 % on the full headline
// frame, -4 % at K0, -8 % on a 1/8 partition (a CTA lives as long as its slowest warp: smaller CTAs return their
// registers sooner and fill the tail of a small grid at a finer grain); flatter 32 x 4 tiles lose 13 % at the oblique
// camera K1 (texture-cache locality wants square tiles), single-warp CTAs lose everywhere.
#ifndef VR_CTA_WARPS
#define VR_CTA_WARPS 4
#endif
#ifndef VR_CTA_WX
#define VR_CTA_WX 2
#endif
#ifndef VR_TP_DEPTH
#define VR_TP_DEPTH 4
#endif
#ifndef VR_TP_MINB
#define VR_TP_MINB 4
#endif
#ifndef VR_TP_SKIP_DEPTH
#define VR_TP_SKIP_DEPTH 3
#endif
#ifndef VR_TP_SKIP_MINB
#define VR_TP_SKIP_MINB 4
#endif
#ifndef VR_NN_DEPTH
#define VR_NN_DEPTH 3     // lab r2: 1.62 ms vs 1.76 ms at depth 2 on the headline volume; 48 registers keep every form free of spills
#endif
#ifndef VR_NN_MINB
#define VR_NN_MINB 5
#endif
#ifndef VR_NN_SKIP_DEPTH
#define VR_NN_SKIP_DEPTH 3
#endif
#ifndef VR_NN_SKIP_MINB
#define VR_NN_SKIP_MINB 5
#endif

}  // namespace

struct vr_context {
    int device = 0;
    int W = 0, H = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float* d_frame = nullptr;            // W*H*4 (+ barrier words)
    uint8_t* d_rgb8 = nullptr;           // W*H*3
    // volume: edge-replicated linear copy (always), layered arrays built on first use
    void* d_vol = nullptr;
    uint64_t vol_bytes = 0;
    int32_t dim[3] = {0, 0, 0};
    int bpv = 0;
    uint32_t pitch = 0;                  // elements
    uint64_t slice = 0;                  // elements
    cudaArray_t d_arr = nullptr;         // source-type layered array (layer = z), point sampling: NEAREST_TEX
    cudaTextureObject_t tex = 0;
    bool arr_failed = false;
    cudaArray_t d_arr2 = nullptr;        // z-pair layered array, Nz+1 layers: TEXPAIR_PIPE
    cudaTextureObject_t tex2 = 0;
    bool arr2_failed = false;
    // per-cell min/max table + empty map for the current min_val
    uint16_t* d_cell_min = nullptr;
    uint16_t* d_cell_max = nullptr;
    uint32_t* d_cell_bits = nullptr;      // empty-cell bit map for `empty_thresh`
    unsigned long long* d_cell_count = nullptr;
    int cell_shift = 0, cells[3] = {0, 0, 0};
    uint64_t ncells = 0, empty_cells = 0;
    bool empty_valid = false;
    int32_t empty_thresh = 0;
    float voxel_size[3] = {1.f, 1.f, 1.f};
    vr_volume_stats stats{};
    bool have_stats = false;
    // state
    float cam[21];
    bool have_cam = false;
    vr_params params;
    float* d_lut = nullptr;              // [0,256): the caller's LUT; [256,512): host-finalised opacity
    bool lut_fast_ok = false;            // every LUT entry finite and >= +0
    float lut_final_host[256];           // what [256,512) of d_lut holds (uploaded only when it changes)
    bool lut_final_valid = false;
    int rank = 0, world = 1, tile_rows = 8;
    // Markstein verification cache: divisor bits -> ok
    std::map<uint32_t, bool> div_ok;
    unsigned int* d_flag = nullptr;
    // fused peer hand-off
    unsigned int* d_done = nullptr;      // CTAs finished (march kernel epilogue): 4 counters, frame_no & 3, so that consecutive
                                         // frames marching concurrently on different streams do not share one
    int done_slot = 0;
    // launch order of the CTA tiles (longest rays first), one table per band, rebuilt on the GPU when its key changes
    struct OrderTable { uint32_t* d = nullptr; size_t cap = 0; cudaEvent_t built = nullptr; std::vector<cudaStream_t> users; std::vector<unsigned char> key; };   // users: streams whose kernels may still read it
    static constexpr int ORDER_TABLES = 12;
    OrderTable order[ORDER_TABLES];
    int order_evict = 0;
    unsigned int* h_peer_error = nullptr;   // mapped pinned word: a barrier wait gave up
    unsigned int* d_peer_error = nullptr;
    // row bands on their own streams: a band's device->host copy overlaps the march of the following bands
    static constexpr int BANDS = 8;      // allocated; VR_BANDS (lab) / band_count() choose how many are used
    cudaStream_t band_stream[BANDS] = {};
    cudaEvent_t band_kdone[BANDS] = {}, band_cdone[BANDS] = {};
    bool bands_ready = false;
    // march-kernel brackets of asynchronous vr_render_peer calls (stats == NULL), read back by vr_peer_kernel_ms
    static constexpr int PEER_EVENTS = 64;
    cudaEvent_t peer_ev[PEER_EVENTS][2] = {};
    uint32_t peer_ev_frame[PEER_EVENTS] = {};
    // frames in flight of the pipelined host path (vr_render_submit / vr_render_wait)
    struct FrameSlot { cudaEvent_t ev0 = nullptr, ev1 = nullptr, done = nullptr; uint32_t launches = 0; int kernel = 0; bool skip = false, pending = false; };
    static constexpr int SLOTS = 2;
    FrameSlot slot[SLOTS];
    uint32_t next_ticket = 1;            // ticket t lives in slot t % SLOTS
    uint32_t slot_ticket[SLOTS] = {0, 0};
};

namespace {

// rows of the compact image: every owned tile padded to tile_rows (keeps the gather regular)
int compact_rows_of(int H, int rank, int world, int tile_rows)
{
    const int tiles = (H + tile_rows - 1) / tile_rows;
    const int owned_tiles = tiles > rank ? (tiles - rank + world - 1) / world : 0;
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

// ---- layered arrays, built from the padded copy on the first frame that needs them -------------------
cudaError_t make_point_texture(cudaArray_t arr, cudaTextureObject_t* tex)
{
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    return cudaCreateTextureObject(tex, &rd, &td, nullptr);
}

// limits of layered 2-D arrays: 32768 x 32768 x 2048 layers
template <typename T>
bool ensure_array_t(vr_context* c)
{
    const int nx = c->dim[0], ny = c->dim[1], nz = c->dim[2];
    if (nz > 2048 || nx > 32768 || ny > 32768) return false;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(8 * (int)sizeof(T), 0, 0, 0, cudaChannelFormatKindUnsigned);
    cudaArray_t arr = nullptr;
    if (cudaMalloc3DArray(&arr, &cd, make_cudaExtent(nx, ny, nz), cudaArrayLayered) != cudaSuccess) { cudaGetLastError(); return false; }
    // interior of the padded copy: voxel (0,0,0) sits at padded (1,1,1); slice stride = pitch * (ny+2) rows
    T* src = static_cast<T*>(c->d_vol) + (c->slice + c->pitch + 1);
    cudaMemcpy3DParms cp = {};
    cp.srcPtr = make_cudaPitchedPtr(src, (size_t)c->pitch * sizeof(T), nx, ny + 2);
    cp.dstArray = arr; cp.extent = make_cudaExtent(nx, ny, nz); cp.kind = cudaMemcpyDeviceToDevice;
    cudaTextureObject_t tex = 0;
    if (cudaMemcpy3DAsync(&cp, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess &&
        make_point_texture(arr, &tex) == cudaSuccess) {
        c->d_arr = arr; c->tex = tex;
        return true;
    }
    cudaFreeArray(arr);
    cudaGetLastError();
    return false;
}

bool ensure_array(vr_context* c)
{
    if (c->tex) return true;
    if (c->arr_failed || !c->d_vol) return false;
    const bool ok = c->bpv == 1 ? ensure_array_t<uint8_t>(c) : ensure_array_t<uint16_t>(c);
    c->arr_failed = !ok;
    return ok;
}

// z-pair array: Nz+1 layers of (z, z+1) words, packed from the padded copy and copied in chunks of layers
// through a bounded (256 MiB) staging buffer
template <typename T, typename W>
bool ensure_zpair_array_t(vr_context* c)
{
    const int nx = c->dim[0], ny = c->dim[1], nz = c->dim[2];
    if (nz + 1 > 2048 || nx > 32768 || ny > 32768) return false;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(8 * (int)sizeof(W), 0, 0, 0, cudaChannelFormatKindUnsigned);
    cudaArray_t arr = nullptr;
    DevBuf stage;
    const uint64_t per_layer = (uint64_t)nx * ny;
    const int chunk = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)nz + 1, (256ull << 20) / (per_layer * sizeof(W))));
    bool ok = cudaMalloc3DArray(&arr, &cd, make_cudaExtent(nx, ny, nz + 1), cudaArrayLayered) == cudaSuccess &&
              stage.alloc(per_layer * (uint64_t)chunk * sizeof(W)) == cudaSuccess;
    for (int l0 = 0; ok && l0 <= nz; l0 += chunk) {
        const int nl = std::min(chunk, nz + 1 - l0);
        vr::zpair_pack_kernel<T, W><<<c->sm_count * 16, 256, 0, c->stream>>>(static_cast<const T*>(c->d_vol), stage.as<W>(), nx, ny,
                                                                             c->pitch, c->slice, l0, nl);
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(stage.p, (size_t)nx * sizeof(W), nx, ny);
        cp.dstArray = arr; cp.dstPos = make_cudaPos(0, 0, l0);
        cp.extent = make_cudaExtent(nx, ny, nl); cp.kind = cudaMemcpyDeviceToDevice;
        ok = cudaGetLastError() == cudaSuccess && cudaMemcpy3DAsync(&cp, c->stream) == cudaSuccess;
    }
    ok = ok && cudaStreamSynchronize(c->stream) == cudaSuccess;
    cudaTextureObject_t tex = 0;
    ok = ok && make_point_texture(arr, &tex) == cudaSuccess;
    if (ok) { c->d_arr2 = arr; c->tex2 = tex; return true; }
    if (arr) cudaFreeArray(arr);
    cudaGetLastError();
    return false;
}

bool ensure_zpair_array(vr_context* c)
{
    if (c->tex2) return true;
    if (c->arr2_failed || !c->d_vol) return false;
    const bool ok = c->bpv == 1 ? ensure_zpair_array_t<uint8_t, uint16_t>(c) : ensure_zpair_array_t<uint16_t, uint32_t>(c);
    c->arr2_failed = !ok;
    return ok;
}

// empty map for the window's lower bound: cell empty <=> cell max <= min_val
// Frames of the pipelined host path that are still in flight read the context's device state (LUTs, empty-cell map,
// launch-order tables, layered arrays): anything that rewrites such state waits for them first.  The tickets stay valid.
int drain_in_flight(vr_context* c)
{
    for (auto& f : c->slot) if (f.pending) VR_CUDA(cudaEventSynchronize(f.done));
    return VR_OK;
}

int ensure_empty_map(vr_context* c, int32_t min_val)
{
    if (!c->d_cell_max) return fail(VR_ERR_NO_VOLUME, "no cell table");
    if (c->empty_valid && c->empty_thresh == min_val) return VR_OK;
    { const int rc = drain_in_flight(c); if (rc != VR_OK) return rc; }
    VR_CUDA(cudaMemsetAsync(c->d_cell_count, 0, sizeof(unsigned long long), c->stream));
    vr::cell_empty_kernel<<<std::max<int>(1, (int)std::min<uint64_t>((c->ncells + 255) / 256, (uint64_t)c->sm_count * 8)), 256, 0, c->stream>>>(
        c->d_cell_max, c->ncells, (unsigned int)min_val, c->d_cell_bits, c->d_cell_count);
    VR_CUDA(cudaGetLastError());
    unsigned long long n = 0;
    VR_CUDA(cudaMemcpyAsync(&n, c->d_cell_count, sizeof n, cudaMemcpyDeviceToHost, c->stream));
    VR_CUDA(cudaStreamSynchronize(c->stream));
    c->empty_cells = n; c->empty_thresh = min_val; c->empty_valid = true;
    return VR_OK;
}

// ---- per-frame plan --------------------------------------------------------------------------------------
enum LoopShape { SHAPE_UNIT = 0, SHAPE_RECIP = 1, SHAPE_MARK = 2, SHAPE_MARK_CAP = 3 };

struct LaunchPlan {
    vr::FrameConsts fc;
    int local_rows;       // rows this rank owns (compact height)
    bool fast;            // the optimised kernels cover this frame
    int form;             // vr::MarchForm
    int shape;            // LoopShape
    int win;              // vr::WinMode
    bool skip;            // empty-space skipping form
    bool lut_final;       // the kernel reads the host-finalised opacity LUT
    int kernel;           // VR_KERNEL_* that will run
};

int make_plan(vr_context* c, int compact, LaunchPlan* plan)
{
    if (!c->d_vol) return fail(VR_ERR_NO_VOLUME, "render: no volume uploaded");
    if (!c->have_cam) return fail(VR_ERR_INVALID, "render: no camera set");
    vr::FrameConsts& fc = plan->fc;
    std::memset(&fc, 0, sizeof fc);
    vr::compute_frame_consts(fc, c->W, c->H, c->dim, c->voxel_size, c->cam, c->params);
    fc.rank = c->rank; fc.world = c->world; fc.tile_rows = c->tile_rows; fc.compact = compact; fc.row0 = 0;
    plan->local_rows = compact_rows_of(c->H, c->rank, c->world, c->tile_rows);

    const vr_params& p = c->params;
    const bool mip = p.is_mip == 1, tf = p.use_tf != 0, swizzled = p.view_top == 1 || p.view_bottom == 1;
    const bool oc = fc.opacity_correction != 0 && !mip;       // the oracle's MIP branch has no opacity correction
    // What the optimised kernels assume: an ordered, non-empty window whose range passes the Markstein check; range
    // tests on float bit patterns need alpha_scale >= +0 and a finite non-negative LUT; correctly rounded tex-coord
    // division without div.rn.  Everything else (including every mode on such a frame) takes the generic loop.
    bool fast = p.max_val > p.min_val && p.alpha_scale >= 0.0f && !std::signbit(p.alpha_scale) && (!tf || c->lut_fast_ok);
    if (fast) {
        bool ok = false;
        int rc = verify_divisor(c, fc.frange, &ok);
        if (rc != VR_OK) return rc;
        fast = ok;
    }
    int tcdiv = fc.tc_div_mode;
    if (fast && tcdiv == vr::DIV_MARKSTEIN) {
        for (int i = 0; i < 3 && fast; ++i) {
            if (vr::is_pow2_float(fc.denom[i])) continue;
            bool ok = false;
            int rc = verify_divisor(c, fc.denom[i], &ok);
            if (rc != VR_OK) return rc;
            fast = ok;
        }
    }
    // opacity correction: with a transfer function the final opacity is a function of the LUT index only and is
    // finalised on the host (vr_set_params); without one it is evaluated per sample and needs a in [0,1]
    plan->lut_final = false;
    if (fast && oc) {
        if (tf) plan->lut_final = true; else fast = p.alpha_scale <= 1.0f;
    }
    if (fast && plan->lut_final) {
        // lut_final[i] = 1 - (1 - lut[i]*alpha)^step_scale, double precision like the oracle (oracle/march_oracle.c:203-208)
        float fin[256];
        for (int i = 0; i < 256 && fast; ++i) {
            float a = p.tf_lut[i] * p.alpha_scale;
            fin[i] = (float)(1.0 - std::pow(1.0 - (double)a, (double)p.step_scale));
            fast = std::isfinite(fin[i]) && fin[i] >= 0.0f && !std::signbit(fin[i]);
        }
        if (fast && !(c->lut_final_valid && std::memcmp(c->lut_final_host, fin, sizeof fin) == 0)) {
            { const int rc = drain_in_flight(c); if (rc != VR_OK) return rc; }
            VR_CUDA(cudaMemcpyAsync(c->d_lut + 256, fin, sizeof fin, cudaMemcpyHostToDevice, c->stream));
            VR_CUDA(cudaStreamSynchronize(c->stream));      // `fin` is on the stack
            std::memcpy(c->lut_final_host, fin, sizeof fin); c->lut_final_valid = true;
        }
    }
    if (!fast) tcdiv = vr::DIV_IEEE;
    fc.tc_div_mode = tcdiv;
    plan->fast = fast;

    // loop shape
    const bool recip = tcdiv == vr::DIV_RECIP_EXACT;
    const bool unit = recip && fc.denom[0] == 1.0f && fc.denom[1] == 1.0f && fc.denom[2] == 1.0f;
    // VolumeRenderer.cs:115 caps the loop at 10000 iterations; a ray cannot take more than
    // |box diagonal| / step + 2 = |vol_size| / step_scale + 2 samples
    const double nmax = std::sqrt((double)c->dim[0] * c->dim[0] + (double)c->dim[1] * c->dim[1] + (double)c->dim[2] * c->dim[2]) /
                        (double)fc.step_scale + 4.0;
    const bool nocap = nmax < 10000.0;
    plan->form = (swizzled || (mip && tf) || plan->lut_final) ? vr::FORM_GENERAL
               : (oc && !tf) ? vr::FORM_GENERAL_OC
               : mip ? vr::FORM_MIP : tf ? vr::FORM_TF : vr::FORM_DVR;
    if (plan->form >= vr::FORM_GENERAL) plan->shape = nocap ? SHAPE_MARK : SHAPE_MARK_CAP;
    else plan->shape = !nocap ? SHAPE_MARK_CAP : unit ? SHAPE_UNIT : recip ? SHAPE_RECIP : SHAPE_MARK;
    plan->win = (p.min_val == 0 && c->have_stats && c->stats.max_value <= p.max_val &&
                 (plan->form == vr::FORM_DVR || plan->form == vr::FORM_TF)) ? vr::WIN_COVERS0 : vr::WIN_CLAMP;

    // kernel
    int want = p.kernel;
    if (!fast) want = VR_KERNEL_DIRECT;
    else if (want != VR_KERNEL_DIRECT) {
        if (fc.filter == VR_FILTER_TRILINEAR) want = ensure_zpair_array(c) ? VR_KERNEL_TEXPAIR_PIPE : VR_KERNEL_DIRECT;
        else want = ensure_array(c) ? VR_KERNEL_NEAREST_TEX : VR_KERNEL_DIRECT;
    }
    plan->kernel = want;

    // empty-space skipping: valid when a sample of value <= min_val contributes exactly nothing
    plan->skip = false;
    if (want != VR_KERNEL_DIRECT && p.empty_skip != VR_SKIP_OFF && p.min_val >= 0 && c->d_cell_max) {
        const float a0 = !tf ? 0.0f : plan->lut_final ? 1.0f /* checked below */ : p.tf_lut[0] * p.alpha_scale;
        bool valid = a0 == 0.0f;
        if (tf && plan->lut_final) valid = (float)(1.0 - std::pow(1.0 - (double)(p.tf_lut[0] * p.alpha_scale), (double)p.step_scale)) == 0.0f;
        if (valid) {
            int rc = ensure_empty_map(c, p.min_val);
            if (rc != VR_OK) return rc;
            // AUTO: the checkpoints cost ~18 % on a frame that has nothing to skip (lab r2: 2.71 -> 3.19 ms) and pay
            // 1.5-2.4x once most cells are empty; switch the form on when at least a quarter of the cells is empty
            plan->skip = p.empty_skip == VR_SKIP_ON ? true : c->empty_cells * 4 >= c->ncells;
        }
    }
    return VR_OK;
}

// ---- launchers -------------------------------------------------------------------------------------------
template <typename T, bool COUNT>
int launch_direct_t(vr_context* c, const LaunchPlan& plan, const vr::DirectArgs& args, dim3 grid, cudaStream_t s)
{
    using namespace vr;
    const dim3 block(DIRECT_BLOCK_W * DIRECT_BLOCK_H);
    // generic loop: run-time filter / MIP / TF / view swizzle / opacity correction, IEEE divisions
    march_direct_kernel<T, VR_FILTER_NEAREST, DIV_IEEE, true, COUNT><<<grid, block, 0, s>>>(plan.fc, args);
    VR_CUDA(cudaGetLastError());
    (void)c;
    return VR_OK;
}

template <typename T, int FORM, bool SKIP>
void launch_texpair_form(const LaunchPlan& plan, const vr::MarchArgs& a, dim3 grid, cudaStream_t s)
{
    using namespace vr;
#define VR_K(TCDIV, WIN, UNIT, NOCAP) march_texpair_kernel<T, TCDIV, WIN, UNIT, NOCAP, FORM, SKIP ? VR_TP_SKIP_DEPTH : VR_TP_DEPTH, SKIP, 8 * (SKIP ? VR_TP_SKIP_MINB : VR_TP_MINB), VR_CTA_WARPS, VR_CTA_WX><<<grid, 32 * VR_CTA_WARPS, SKIP ? (size_t)a.cell_words * 4 : 0, s>>>(plan.fc, a)
    if constexpr (FORM >= FORM_GENERAL) {
        if (plan.shape == SHAPE_MARK) VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, true); else VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, false);
    } else {
        if constexpr (FORM == FORM_DVR || FORM == FORM_TF) {
            if (plan.win == WIN_COVERS0) {
                switch (plan.shape) {
                    case SHAPE_UNIT:  VR_K(DIV_RECIP_EXACT, WIN_COVERS0, true, true); return;
                    case SHAPE_RECIP: VR_K(DIV_RECIP_EXACT, WIN_COVERS0, false, true); return;
                    case SHAPE_MARK:  VR_K(DIV_MARKSTEIN, WIN_COVERS0, false, true); return;
                    default:          VR_K(DIV_MARKSTEIN, WIN_COVERS0, false, false); return;
                }
            }
        }
        switch (plan.shape) {
            case SHAPE_UNIT:  VR_K(DIV_RECIP_EXACT, WIN_CLAMP, true, true); return;
            case SHAPE_RECIP: VR_K(DIV_RECIP_EXACT, WIN_CLAMP, false, true); return;
            case SHAPE_MARK:  VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, true); return;
            default:          VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, false); return;
        }
    }
#undef VR_K
}

template <int FORM, bool SKIP>
void launch_nearest_form(const LaunchPlan& plan, const vr::MarchArgs& a, dim3 grid, cudaStream_t s)
{
    using namespace vr;
#define VR_K(TCDIV, WIN, UNIT, NOCAP) march_nearest_kernel<TCDIV, WIN, UNIT, NOCAP, FORM, SKIP ? VR_NN_SKIP_DEPTH : VR_NN_DEPTH, SKIP, 8 * (SKIP ? VR_NN_SKIP_MINB : (FORM >= vr::FORM_GENERAL ? 5 : VR_NN_MINB)), VR_CTA_WARPS, VR_CTA_WX><<<grid, 32 * VR_CTA_WARPS, SKIP ? (size_t)a.cell_words * 4 : 0, s>>>(plan.fc, a)
    if constexpr (FORM >= FORM_GENERAL) {
        if (plan.shape == SHAPE_MARK) VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, true); else VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, false);
    } else {
        if constexpr (FORM == FORM_DVR || FORM == FORM_TF) {
            if (plan.win == WIN_COVERS0) {
                switch (plan.shape) {
                    case SHAPE_UNIT:  VR_K(DIV_RECIP_EXACT, WIN_COVERS0, true, true); return;
                    case SHAPE_RECIP: VR_K(DIV_RECIP_EXACT, WIN_COVERS0, false, true); return;
                    case SHAPE_MARK:  VR_K(DIV_MARKSTEIN, WIN_COVERS0, false, true); return;
                    default:          VR_K(DIV_MARKSTEIN, WIN_COVERS0, false, false); return;
                }
            }
        }
        switch (plan.shape) {
            case SHAPE_UNIT:  VR_K(DIV_RECIP_EXACT, WIN_CLAMP, true, true); return;
            case SHAPE_RECIP: VR_K(DIV_RECIP_EXACT, WIN_CLAMP, false, true); return;
            case SHAPE_MARK:  VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, true); return;
            default:          VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, false); return;
        }
    }
#undef VR_K
}

template <typename T, bool SKIP>
void launch_texpair_ts(const LaunchPlan& plan, const vr::MarchArgs& a, dim3 grid, cudaStream_t s)
{
    switch (plan.form) {
        case vr::FORM_DVR:        launch_texpair_form<T, vr::FORM_DVR, SKIP>(plan, a, grid, s); break;
        case vr::FORM_TF:         launch_texpair_form<T, vr::FORM_TF, SKIP>(plan, a, grid, s); break;
        case vr::FORM_MIP:        launch_texpair_form<T, vr::FORM_MIP, SKIP>(plan, a, grid, s); break;
        case vr::FORM_GENERAL:    launch_texpair_form<T, vr::FORM_GENERAL, SKIP>(plan, a, grid, s); break;
        default:                  launch_texpair_form<T, vr::FORM_GENERAL_OC, SKIP>(plan, a, grid, s); break;
    }
}

template <bool SKIP>
void launch_nearest_s(const LaunchPlan& plan, const vr::MarchArgs& a, dim3 grid, cudaStream_t s)
{
    switch (plan.form) {
        case vr::FORM_DVR:        launch_nearest_form<vr::FORM_DVR, SKIP>(plan, a, grid, s); break;
        case vr::FORM_TF:         launch_nearest_form<vr::FORM_TF, SKIP>(plan, a, grid, s); break;
        case vr::FORM_MIP:        launch_nearest_form<vr::FORM_MIP, SKIP>(plan, a, grid, s); break;
        case vr::FORM_GENERAL:    launch_nearest_form<vr::FORM_GENERAL, SKIP>(plan, a, grid, s); break;
        default:                  launch_nearest_form<vr::FORM_GENERAL_OC, SKIP>(plan, a, grid, s); break;
    }
}

// Launch order of the CTA tiles: tiles sorted by the estimated length of their rays, longest first (LPT), so that what
// runs on the draining machine at the end of the grid are the short rays.  Lab r2 (B200, headline frame): 2.47 -> 2.36 ms
// on the full frame, 1.10 -> 0.94 ms at the oblique camera K1, 0.443 -> 0.355 ms on a 1/8 partition.  Only the ORDER in
// which the hardware starts the tiles changes; every pixel is computed by the same code from the same inputs.  The table
// is built on the GPU by one small kernel in the frame's stream (cta_order_kernel: chord of each tile's centre and two corner
// rays through the box, counting sort over 1024 length classes) and only when its key -- camera, box, partition, band --
// changes; one table per band.  VR_LPT=0 turns it off.
const uint32_t* cta_order_for(vr_context* c, const LaunchPlan& plan, dim3 grid, int px_w, int px_h, int row0, int row_end, cudaStream_t s)
{
    static const int mode = [] { const char* e = std::getenv("VR_LPT"); return e ? std::atoi(e) : 1; }();
    const size_t n = (size_t)grid.x * grid.y;
    if (mode == 0 || n < 2 || n > (1u << 20) || grid.x > 65535u || grid.y > 65535u) return nullptr;
    const vr::FrameConsts& fc = plan.fc;
    struct Key { float cam[21]; float pmin[3], pmax[3]; int v[11]; } key;
    std::memset(&key, 0, sizeof key);
    std::memcpy(key.cam, fc.cam, sizeof key.cam);
    std::memcpy(key.pmin, fc.pmin, sizeof key.pmin); std::memcpy(key.pmax, fc.pmax, sizeof key.pmax);
    const int kv[11] = {fc.W, fc.H, fc.rank, fc.world, fc.tile_rows, row0, row_end, (int)grid.x, (int)grid.y, px_w, px_h};
    std::memcpy(key.v, kv, sizeof kv);
    // the band's table: same (row0, row_end, partition) slot if there is one, else a free one, else evict round-robin
    vr_context::OrderTable* t = nullptr;
    for (auto& o : c->order)
        if (o.key.size() == sizeof key && std::memcmp(o.key.data() + offsetof(Key, v), key.v, sizeof key.v) == 0) { t = &o; break; }
    if (!t) for (auto& o : c->order) if (o.key.empty()) { t = &o; break; }
    if (!t) { t = &c->order[c->order_evict]; c->order_evict = (c->order_evict + 1) % vr_context::ORDER_TABLES; }
    if (t->key.size() == sizeof key && std::memcmp(t->key.data(), &key, sizeof key) == 0 && t->d) {
        if (std::find(t->users.begin(), t->users.end(), s) == t->users.end()) {
            // another stream than the one that built the table: order this stream's kernels behind the build
            if (cudaStreamWaitEvent(s, t->built, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            t->users.push_back(s);
        }
        return t->d;
    }
    // rebuild.  Kernels of earlier frames on OTHER streams may still read the table: wait for those streams first
    for (cudaStream_t u : t->users) if (u != s && cudaStreamSynchronize(u) != cudaSuccess) return nullptr;
    if (2 * n > t->cap) {
        if (t->d) { cudaDeviceSynchronize(); cudaFree(t->d); }
        t->d = nullptr; t->cap = 0; t->key.clear();
        if (cudaMalloc(&t->d, 2 * n * sizeof(uint32_t)) != cudaSuccess) { t->d = nullptr; cudaGetLastError(); return nullptr; }
        t->cap = 2 * n;
    }
    vr::cta_order_kernel<<<1, 1024, 0, s>>>(fc, row0, row_end, px_w, px_h, (int)grid.x, (int)grid.y, t->d, t->d + n);
    if (cudaGetLastError() != cudaSuccess) { t->key.clear(); return nullptr; }
    if (!t->built && cudaEventCreateWithFlags(&t->built, cudaEventDisableTiming) != cudaSuccess) { t->built = nullptr; t->key.clear(); cudaGetLastError(); return nullptr; }
    if (cudaEventRecord(t->built, s) != cudaSuccess) { t->key.clear(); cudaGetLastError(); return nullptr; }
    t->key.assign(reinterpret_cast<unsigned char*>(&key), reinterpret_cast<unsigned char*>(&key) + sizeof key);
    t->users.assign(1, s);
    return t->d;
}

#ifdef VR_LAB
// development builds only (tools/lab/build_lab.sh): pipeline depth / occupancy variants of the DVR form, chosen per
// launch through the environment: VR_LAB_TP="depth,minb"; returns -1 when the frame / variant is not covered
template <typename T, int DEPTH, int MINW, bool SKIP, int CTAW, int WX = (CTAW < 4 ? CTAW : 4)>
int lab_tp(vr_context* c, const LaunchPlan& plan, const vr::MarchArgs& a, int W, int row0, int row_end, cudaStream_t s)
{
    using namespace vr;
    typedef CtaShape<CTAW, WX> S;
    const dim3 grid((W + S::PX - 1) / S::PX, (row_end - row0 + S::PY - 1) / S::PY);
    MarchArgs b = a; b.grid_ctas = grid.x * grid.y;
    b.cta_order = cta_order_for(c, plan, grid, S::PX, S::PY, row0, row_end, s);   // a table for this tile shape
#define VR_K(TCDIV, WIN, UNIT, NOCAP) march_texpair_kernel<T, TCDIV, WIN, UNIT, NOCAP, FORM_DVR, DEPTH, SKIP, MINW, CTAW, WX><<<grid, S::THREADS, SKIP ? (size_t)a.cell_words * 4 : 0, s>>>(plan.fc, b)
    if (plan.shape == SHAPE_UNIT && plan.win == WIN_COVERS0) { VR_K(DIV_RECIP_EXACT, WIN_COVERS0, true, true); return 0; }
    if (plan.shape == SHAPE_UNIT && plan.win == WIN_CLAMP)   { VR_K(DIV_RECIP_EXACT, WIN_CLAMP, true, true); return 0; }
    if (plan.shape == SHAPE_MARK && plan.win == WIN_CLAMP)   { VR_K(DIV_MARKSTEIN, WIN_CLAMP, false, true); return 0; }
#undef VR_K
    return -1;
}
template <typename T, bool SKIP>
int lab_tp_s(vr_context* c, const LaunchPlan& plan, const vr::MarchArgs& a, int W, int row0, int row_end, cudaStream_t s, int depth, int minb, int ctaw)
{
    // minb = resident warps per SM (register budget), ctaw = warps per CTA
    const int key = depth * 10000 + minb * 100 + ctaw;
    switch (key) {
        case 24808: return lab_tp<T, 2, 48, SKIP, 8>(c, plan, a, W, row0, row_end, s);
        case 33208: return lab_tp<T, 3, 32, SKIP, 8>(c, plan, a, W, row0, row_end, s);
        case 34008: return lab_tp<T, 3, 40, SKIP, 8>(c, plan, a, W, row0, row_end, s);
        case 43208: return lab_tp<T, 4, 32, SKIP, 8>(c, plan, a, W, row0, row_end, s);
        case 43204: return lab_tp<T, 4, 32, SKIP, 4>(c, plan, a, W, row0, row_end, s);
        case 43202: return lab_tp<T, 4, 32, SKIP, 2>(c, plan, a, W, row0, row_end, s);
        case 43201: return lab_tp<T, 4, 32, SKIP, 1>(c, plan, a, W, row0, row_end, s);
        case 33604: return lab_tp<T, 3, 36, SKIP, 4>(c, plan, a, W, row0, row_end, s);
        case 33602: return lab_tp<T, 3, 36, SKIP, 2>(c, plan, a, W, row0, row_end, s);
        case 43604: return lab_tp<T, 4, 36, SKIP, 4>(c, plan, a, W, row0, row_end, s);
        case 43602: return lab_tp<T, 4, 36, SKIP, 2>(c, plan, a, W, row0, row_end, s);
        case 34004: return lab_tp<T, 3, 40, SKIP, 4>(c, plan, a, W, row0, row_end, s);
        case 34002: return lab_tp<T, 3, 40, SKIP, 2>(c, plan, a, W, row0, row_end, s);
        // squarer tiles: ctaw code 42 = 4 warps as 16x8 px, 21 = 2 warps as 8x8, 82 = 8 warps as 16x16, 41 = 4 warps as 8x16
        case 43242: return lab_tp<T, 4, 32, SKIP, 4, 2>(c, plan, a, W, row0, row_end, s);
        case 43221: return lab_tp<T, 4, 32, SKIP, 2, 1>(c, plan, a, W, row0, row_end, s);
        case 43282: return lab_tp<T, 4, 32, SKIP, 8, 2>(c, plan, a, W, row0, row_end, s);
        case 43241: return lab_tp<T, 4, 32, SKIP, 4, 1>(c, plan, a, W, row0, row_end, s);
        case 33642: return lab_tp<T, 3, 36, SKIP, 4, 2>(c, plan, a, W, row0, row_end, s);
        case 42442: return lab_tp<T, 4, 24, SKIP, 4, 2>(c, plan, a, W, row0, row_end, s);
        case 42882: return lab_tp<T, 4, 28, SKIP, 4, 2>(c, plan, a, W, row0, row_end, s);   // 70 registers: 7 CTAs of 4 warps, 52.8 instr/sample
        case 32442: return lab_tp<T, 3, 24, SKIP, 4, 2>(c, plan, a, W, row0, row_end, s);
        default: return -1;
    }
}
int lab_launch_texpair(vr_context* c, const LaunchPlan& plan, const vr::MarchArgs& a, int W, int row0, int row_end, cudaStream_t s, int bpv)
{
    const char* e = std::getenv("VR_LAB_TP");
    int depth = 0, minb = 0, ctaw = 8;
    if (!e || std::sscanf(e, "%d,%d,%d", &depth, &minb, &ctaw) < 2 || plan.form != vr::FORM_DVR) return -1;
    if (bpv == 2) return plan.skip ? lab_tp_s<uint16_t, true>(c, plan, a, W, row0, row_end, s, depth, minb, ctaw)
                                   : lab_tp_s<uint16_t, false>(c, plan, a, W, row0, row_end, s, depth, minb, ctaw);
    return plan.skip ? lab_tp_s<uint8_t, true>(c, plan, a, W, row0, row_end, s, depth, minb, ctaw)
                     : lab_tp_s<uint8_t, false>(c, plan, a, W, row0, row_end, s, depth, minb, ctaw);
}
#endif

// the march.  `row0`/`row_end` select a band of this rank's local rows (compact row space); `peer_arrive` != null
// fuses the hand-off signal into the kernel epilogue (returns *signalled = false when the kernel that ran cannot)
int launch_march(vr_context* c, LaunchPlan& plan, float* d_out, int row0, int row_end, cudaStream_t s,
                 unsigned int* peer_arrive, bool* signalled)
{
    if (signalled) *signalled = false;
    if (row_end <= row0) return VR_OK;                      // this rank owns no row of the band (tiles < world)
    plan.fc.row0 = row0;
    if (plan.kernel == VR_KERNEL_DIRECT) {
        const dim3 grid((c->W + vr::DIRECT_BLOCK_W - 1) / vr::DIRECT_BLOCK_W, (row_end - row0 + vr::DIRECT_BLOCK_H - 1) / vr::DIRECT_BLOCK_H);
        vr::DirectArgs args{};
        args.vol = c->d_vol; args.pitch = c->pitch; args.slice = c->slice;
        args.tf_lut = c->d_lut; args.out = d_out; args.local_rows = row_end;
        return c->bpv == 1 ? launch_direct_t<uint8_t, false>(c, plan, args, grid, s)
                           : launch_direct_t<uint16_t, false>(c, plan, args, grid, s);
    }
    typedef vr::CtaShape<VR_CTA_WARPS, VR_CTA_WX> Shape;
    const dim3 grid((c->W + Shape::PX - 1) / Shape::PX, (row_end - row0 + Shape::PY - 1) / Shape::PY);
    vr::MarchArgs a{};
    a.out = d_out; a.local_rows = row_end;
    a.tf_lut = plan.lut_final ? c->d_lut + 256 : c->d_lut;
    a.lut_final = plan.lut_final ? 1 : 0;
    a.cell_bits = c->d_cell_bits; a.cell_words = (int)((c->ncells + 31) / 32); a.cell_shift = c->cell_shift;
    a.cell_nx = c->cells[0]; a.cell_nxy = c->cells[0] * c->cells[1];
    // checkpoint period in passes of the unrolled loop (lab r2, C4 window [1000,3000] K2, ms): 1: 2.41, 2: 2.00, 4: 1.77,
    // 8: 1.69, 16: 1.72 -- what a checkpoint costs is not its ~25 instructions but the lanes that idle while others leap
    static const int check_every = [] { const char* e = std::getenv("VR_SKIP_CHECK"); const int v = e ? std::atoi(e) : 8; return (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) ? v : 8; }();
    a.skip_check_mask = check_every - 1;
    a.done_counter = c->d_done + (c->done_slot & 3); a.peer_arrive = peer_arrive; a.grid_ctas = grid.x * grid.y;
    a.cta_order = cta_order_for(c, plan, grid, Shape::PX, Shape::PY, row0, row_end, s);
    if (signalled) *signalled = peer_arrive != nullptr;
    if (plan.kernel == VR_KERNEL_TEXPAIR_PIPE) {
        a.tex = c->tex2;
#ifdef VR_LAB
        if (lab_launch_texpair(c, plan, a, c->W, row0, row_end, s, c->bpv) == 0) { VR_CUDA(cudaGetLastError()); return VR_OK; }
#endif
        if (c->bpv == 2) { if (plan.skip) launch_texpair_ts<uint16_t, true>(plan, a, grid, s); else launch_texpair_ts<uint16_t, false>(plan, a, grid, s); }
        else             { if (plan.skip) launch_texpair_ts<uint8_t, true>(plan, a, grid, s);  else launch_texpair_ts<uint8_t, false>(plan, a, grid, s); }
    } else {
        a.tex = c->tex;
        if (plan.skip) launch_nearest_s<true>(plan, a, grid, s); else launch_nearest_s<false>(plan, a, grid, s);
    }
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

void fill_stats(vr_render_stats* stats, float ms, uint32_t launches, const LaunchPlan& plan)
{
    if (!stats) return;
    stats->kernel_ms = ms; stats->kernel_launches = launches; stats->kernel_used = (uint32_t)plan.kernel;
    stats->skip_used = plan.skip ? 1u : 0u;
}

int render_common(vr_context* c, float* d_out, int compact, cudaStream_t s, vr_render_stats* stats)
{
    LaunchPlan plan;
    int rc = make_plan(c, compact, &plan);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaEventRecord(c->ev0, s));
    rc = launch_march(c, plan, d_out, 0, plan.local_rows, s, nullptr, nullptr);
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaEventRecord(c->ev1, s));
    VR_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    VR_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    fill_stats(stats, ms, plan.local_rows > 0 ? 1u : 0u, plan);
    return VR_OK;
}

void release_volume(vr_context* c)
{
    if (c->tex) { cudaDestroyTextureObject(c->tex); c->tex = 0; }
    if (c->d_arr) { cudaFreeArray(c->d_arr); c->d_arr = nullptr; }
    if (c->tex2) { cudaDestroyTextureObject(c->tex2); c->tex2 = 0; }
    if (c->d_arr2) { cudaFreeArray(c->d_arr2); c->d_arr2 = nullptr; }
    if (c->d_vol) { cudaFree(c->d_vol); c->d_vol = nullptr; }
    if (c->d_cell_min) { cudaFree(c->d_cell_min); c->d_cell_min = nullptr; }
    if (c->d_cell_max) { cudaFree(c->d_cell_max); c->d_cell_max = nullptr; }
    if (c->d_cell_bits) { cudaFree(c->d_cell_bits); c->d_cell_bits = nullptr; }
    c->arr_failed = c->arr2_failed = false;
    c->empty_valid = false; c->ncells = 0; c->vol_bytes = 0;
    c->have_stats = false;
}

// Volume ingest from a device-resident x-fastest source (RendererCore.cpp:360-419 on the GPU):
//   pass 1  pad_cells_kernel    ONE read of the source: edge-replicated linear copy + per-cell min/max table + min/max
//                               (source rows 16-byte aligned; otherwise pad_minmax_kernel, and cell_minmax_kernel as pass 3)
//   pass 2  histogram_kernel    (needs the max of pass 1 for the 16-bit binning, RendererCore.cpp:386-398)
// The layered arrays are built later, from the padded copy, by the first frame that needs them.  Everything is
// allocated into locals first; the context is only touched once nothing can fail any more.
// Peak footprint during the call: source + padded copy (+ the previous volume until the swap).
template <typename T>
int ingest_from_device(vr_context* c, const T* d_src, const uint64_t dims[3])
{
    const int nx = (int)dims[0], ny = (int)dims[1], nz = (int)dims[2];
    const uint64_t n = (uint64_t)nx * ny * nz;
    { const int rc = drain_in_flight(c); if (rc != VR_OK) return rc; }
    DevBuf mm, bins, hlut, padded, cmin, cmax, cempty, cell32;
    VR_CUDA(mm.alloc(2 * sizeof(unsigned int)));
    VR_CUDA(bins.alloc(256 * sizeof(unsigned long long)));
    const unsigned int init[2] = {0xffffffffu, 0u};
    VR_CUDA(cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    VR_CUDA(cudaMemsetAsync(bins.p, 0, 256 * sizeof(unsigned long long), c->stream));

    // cell table geometry: smallest cell side >= 8 voxels with at most 65536 cells (the empty map then stays resident
    // in L1 / shared memory next to the texture working set)
    int shift = 3;
    auto cells_of = [&](int sh, int a) { return (int)(dims[a] >> sh) + 1; };
    while ((uint64_t)cells_of(shift, 0) * cells_of(shift, 1) * cells_of(shift, 2) > 65536ull) ++shift;
    const int cnx = cells_of(shift, 0), cny = cells_of(shift, 1), cnz = cells_of(shift, 2);
    const uint64_t ncells = (uint64_t)cnx * cny * cnz;
    VR_CUDA(cmin.alloc(ncells * sizeof(uint16_t)));
    VR_CUDA(cmax.alloc(ncells * sizeof(uint16_t)));
    VR_CUDA(cempty.alloc(((ncells + 31) / 32) * sizeof(uint32_t)));

    // pass 1: padded copy + min/max (+ the cell table when the source rows are 16-byte aligned: ONE read of the source)
    const uint32_t pitch = (uint32_t)(round_up((uint64_t)(nx + 2) * sizeof(T), 16) / sizeof(T));
    const uint64_t slice = (uint64_t)pitch * (uint64_t)(ny + 2);
    const uint64_t bytes = slice * (uint64_t)(nz + 2) * sizeof(T) + 256;
    VR_CUDA(padded.alloc(bytes));
    VR_CUDA(cudaMemsetAsync(static_cast<char*>(padded.p) + bytes - 256, 0, 256, c->stream));
    static const bool no_fused = [] { const char* e = std::getenv("VR_INGEST_FUSED"); return e && std::atoi(e) == 0; }();
    const bool fused = !no_fused && nx % (16 / (int)sizeof(T)) == 0 && (reinterpret_cast<uintptr_t>(d_src) & 15) == 0;
    if (fused) {
        VR_CUDA(cell32.alloc(2 * ncells * sizeof(unsigned int)));
        unsigned int* gmin = cell32.as<unsigned int>(), *gmax = gmin + ncells;
        VR_CUDA(cudaMemsetAsync(gmin, 0xff, ncells * sizeof(unsigned int), c->stream));
        VR_CUDA(cudaMemsetAsync(gmax, 0, ncells * sizeof(unsigned int), c->stream));
        const uint64_t passes = (uint64_t)((ny + 2 + 7) / 8) * (uint64_t)(nz + 2);
        // grid = exactly what is resident at once: the kernel is grid-stride over equal passes, a partial second round of
        // CTAs would leave most of the machine idle at the end
        const size_t cells_smem = 4 * (size_t)cnx * sizeof(unsigned int);
        int per_sm = 0;
        VR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vr::pad_cells_kernel<T>, 256, cells_smem));
        if (per_sm < 1) per_sm = 1;
        vr::pad_cells_kernel<T><<<(unsigned)std::min<uint64_t>(passes, (uint64_t)c->sm_count * per_sm), 256, cells_smem, c->stream>>>(
            d_src, padded.as<T>(), nx, ny, nz, pitch, shift, cnx, cny, gmin, gmax);
        VR_CUDA(cudaGetLastError());
        vr::cell_table_finish_kernel<<<(unsigned)std::min<uint64_t>((ncells + 255) / 256, 64), 256, 0, c->stream>>>(
            gmin, gmax, ncells, cmin.as<uint16_t>(), cmax.as<uint16_t>(), mm.as<unsigned int>(), mm.as<unsigned int>() + 1);
        VR_CUDA(cudaGetLastError());
    } else {
        vr::pad_minmax_kernel<T><<<c->sm_count * 16, 256, 0, c->stream>>>(d_src, padded.as<T>(), nx, ny, nz, pitch,
                                                                          mm.as<unsigned int>(), mm.as<unsigned int>() + 1);
        VR_CUDA(cudaGetLastError());
    }
    unsigned int mmh[2] = {0, 255};
    VR_CUDA(cudaMemcpyAsync(mmh, mm.p, sizeof mmh, cudaMemcpyDeviceToHost, c->stream));
    VR_CUDA(cudaStreamSynchronize(c->stream));

    // pass 2: histogram
    {
        size_t smem = (vr::HIST_THREADS / 32) * 256 * sizeof(unsigned int);
        int lut_entries = 0;
        if (sizeof(T) == 2) {
            VR_CUDA(hlut.alloc(65536));
            vr::histogram_lut_kernel<<<65536 / 256, 256, 0, c->stream>>>((float)(int)mmh[1], hlut.as<uint8_t>());
            lut_entries = (int)std::min<unsigned int>(65536u, (mmh[1] + 16u) & ~15u);      // values never exceed the dataset max
            smem += (size_t)lut_entries;
            VR_CUDA(cudaFuncSetAttribute(vr::histogram_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        int per_sm = 0;
        VR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vr::histogram_kernel<T>, vr::HIST_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        vr::histogram_kernel<T><<<c->sm_count * per_sm, vr::HIST_THREADS, smem, c->stream>>>(d_src, n, hlut.as<uint8_t>(), lut_entries, bins.as<unsigned long long>());
        VR_CUDA(cudaGetLastError());
    }
    unsigned long long hbins[256];
    VR_CUDA(cudaMemcpyAsync(hbins, bins.p, sizeof hbins, cudaMemcpyDeviceToHost, c->stream));

    // pass 3 (only when pass 1 could not build it): cell table from the padded copy
    if (!fused) {
        vr::cell_minmax_kernel<T><<<(unsigned)std::min<uint64_t>(ncells, (uint64_t)c->sm_count * 32), 128, 0, c->stream>>>(
            padded.as<T>(), pitch, slice, nx, ny, nz, shift, cnx, cny, cnz, cmin.as<uint16_t>(), cmax.as<uint16_t>());
        VR_CUDA(cudaGetLastError());
    }
    VR_CUDA(cudaStreamSynchronize(c->stream));          // also completes the D2H copy into `hbins`

    // nothing can fail from here: swap the new volume in
    if (!c->d_cell_count) VR_CUDA(cudaMalloc(&c->d_cell_count, sizeof(unsigned long long)));
    release_volume(c);
    c->d_vol = padded.release(); c->vol_bytes = bytes;
    c->dim[0] = nx; c->dim[1] = ny; c->dim[2] = nz;
    c->bpv = (int)sizeof(T); c->pitch = pitch; c->slice = slice;
    c->d_cell_min = static_cast<uint16_t*>(cmin.release());
    c->d_cell_max = static_cast<uint16_t*>(cmax.release());
    c->d_cell_bits = static_cast<uint32_t*>(cempty.release());
    c->cell_shift = shift; c->cells[0] = cnx; c->cells[1] = cny; c->cells[2] = cnz; c->ncells = ncells;

    // histogram normalisation exactly as RendererCore.cpp:361,386-405: float bins that are
    // incremented one by one saturate at 2^24; max_value starts at the 16-bit dataset max
    // (or -1 for 8-bit data) and is then raised by the bin counts.
    vr_volume_stats& st = c->stats;
    if (sizeof(T) == 2) { st.min_value = (int)mmh[0]; st.max_value = (int)mmh[1]; }
    else { st.min_value = 0; st.max_value = 255; }
    int max_value = sizeof(T) == 2 ? (int)mmh[1] : -1;
    float hist[256];
    for (int i = 0; i < 256; ++i) {
        const unsigned long long cnt = hbins[i] > 16777216ull ? 16777216ull : hbins[i];
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
    p->empty_skip = VR_SKIP_AUTO;
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
    if (e == cudaSuccess) e = cudaMalloc(&c->d_lut, 512 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_flag, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->d_lut, 0, 512 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_done, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(c->d_done, 0, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_peer_error, sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *c->h_peer_error = 0; e = cudaHostGetDevicePointer(&c->d_peer_error, c->h_peer_error, 0); }
    if (e != cudaSuccess) { int rc = cuda_fail(e, "vr_create"); vr_destroy(c); return rc; }
    *out = c;
    return VR_OK;
}

void vr_destroy(vr_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();             // frames in flight, asynchronous peer frames
    release_volume(c);
    if (c->d_cell_count) cudaFree(c->d_cell_count);
    if (c->d_done) cudaFree(c->d_done);
    for (auto& o : c->order) { if (o.d) cudaFree(o.d); if (o.built) cudaEventDestroy(o.built); }
    if (c->h_peer_error) cudaFreeHost(c->h_peer_error);
    if (c->d_frame) cudaFree(c->d_frame);
    if (c->d_rgb8) cudaFree(c->d_rgb8);
    if (c->d_lut) cudaFree(c->d_lut);
    if (c->d_flag) cudaFree(c->d_flag);
    for (int b = 0; b < vr_context::BANDS; ++b) {
        if (c->band_stream[b]) cudaStreamDestroy(c->band_stream[b]);
        if (c->band_kdone[b]) cudaEventDestroy(c->band_kdone[b]);
        if (c->band_cdone[b]) cudaEventDestroy(c->band_cdone[b]);
    }
    for (auto& pe : c->peer_ev) { if (pe[0]) cudaEventDestroy(pe[0]); if (pe[1]) cudaEventDestroy(pe[1]); }
    for (auto& f : c->slot) {
        if (f.pending) cudaEventSynchronize(f.done);
        if (f.ev0) cudaEventDestroy(f.ev0);
        if (f.ev1) cudaEventDestroy(f.ev1);
        if (f.done) cudaEventDestroy(f.done);
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
    { const int rc = drain_in_flight(c); if (rc != VR_OK) return rc; }      // frames in flight read and write the old image
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

int vr_cell_table_get(vr_context* c, int* shift, int cells[3], uint16_t* mins, uint16_t* maxs, uint64_t* empty_cells)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_cell_table_get: null context");
    if (!c->d_cell_max) return fail(VR_ERR_NO_VOLUME, "vr_cell_table_get: no volume uploaded");
    VR_CUDA(cudaSetDevice(c->device));
    if (shift) *shift = c->cell_shift;
    if (cells) for (int i = 0; i < 3; ++i) cells[i] = c->cells[i];
    if (mins) VR_CUDA(cudaMemcpy(mins, c->d_cell_min, c->ncells * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    if (maxs) VR_CUDA(cudaMemcpy(maxs, c->d_cell_max, c->ncells * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    if (empty_cells) {
        *empty_cells = 0;
        if (c->params.min_val >= 0) {
            int rc = ensure_empty_map(c, c->params.min_val);
            if (rc != VR_OK) return rc;
            *empty_cells = c->empty_cells;
        }
    }
    return VR_OK;
}

int vr_memory_info_get(vr_context* c, vr_memory_info* out)
{
    if (!c || !out) return fail(VR_ERR_INVALID, "vr_memory_info_get: null");
    std::memset(out, 0, sizeof *out);
    out->frame_bytes = frame_alloc_bytes(c->W, c->H);
    if (!c->d_vol) return VR_OK;
    const uint64_t nvox = (uint64_t)c->dim[0] * c->dim[1] * c->dim[2];
    out->linear_bytes = c->vol_bytes;
    out->array_bytes = c->d_arr ? nvox * (uint64_t)c->bpv : 0;
    out->zpair_array_bytes = c->d_arr2 ? (uint64_t)c->dim[0] * c->dim[1] * (uint64_t)(c->dim[2] + 1) * 2ull * (uint64_t)c->bpv : 0;
    out->cell_table_bytes = c->ncells * 4ull + ((c->ncells + 31) / 32) * 4ull;
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
    if (p->kernel != VR_KERNEL_AUTO && p->kernel != VR_KERNEL_DIRECT && p->kernel != VR_KERNEL_TEXPAIR_PIPE && p->kernel != VR_KERNEL_NEAREST_TEX)
        return fail(VR_ERR_INVALID, "vr_set_params: unknown kernel (AUTO, DIRECT, TEXPAIR_PIPE, NEAREST_TEX)");
    if (p->empty_skip < VR_SKIP_AUTO || p->empty_skip > VR_SKIP_OFF)
        return fail(VR_ERR_INVALID, "vr_set_params: unknown empty_skip mode");
    if (std::memcmp(p, &c->params, sizeof *p) == 0) return VR_OK;          // unchanged (a per-frame caller): nothing to do
    if (p->use_tf) {
        { const int rc = drain_in_flight(c); if (rc != VR_OK) return rc; }
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
    if (rank != c->rank || world != c->world || tile_rows != c->tile_rows) { const int rc = drain_in_flight(c); if (rc != VR_OK) return rc; }
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
    if (!d_rgba && (c->slot[0].pending || c->slot[1].pending))
        return fail(VR_ERR_INVALID, "vr_render_device: the context's own frame is in use by frames submitted with vr_render_submit");
    if (!d_rgba) { d_rgba = c->d_frame; compact = 0; }
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    int rc = render_common(c, d_rgba, compact ? 1 : 0, s, stats);
    if (rc != VR_OK) return rc;
    if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

// how many bands a banded render uses (<= vr_context::BANDS): the copy of the LAST band is the part of the transfer that
// is not hidden behind the march, so large frames want more bands (1080p on one GPU: 2 / 4 / 6 / 8 bands ->
// 2.80 / 2.66 / 2.60 / 2.58 ms end to end); VR_BANDS overrides
static int band_count(size_t owned_px)
{
    static const int forced = [] { const char* e = std::getenv("VR_BANDS"); return e ? std::atoi(e) : 0; }();
    const int n = forced > 0 ? forced : (owned_px >= ((size_t)1 << 20) ? 8 : 4);
    return n < 1 ? 1 : (n > vr_context::BANDS ? vr_context::BANDS : n);
}

static int ensure_bands(vr_context* c)
{
    if (c->bands_ready) return VR_OK;
    // (stream priorities for the earlier bands were measured: no effect, CTAs are dispatched in launch order anyway)
    for (int b = 0; b < vr_context::BANDS; ++b) {
        VR_CUDA(cudaStreamCreateWithFlags(&c->band_stream[b], cudaStreamNonBlocking));
        VR_CUDA(cudaEventCreateWithFlags(&c->band_kdone[b], cudaEventDisableTiming));
        VR_CUDA(cudaEventCreateWithFlags(&c->band_cdone[b], cudaEventDisableTiming));
    }
    c->bands_ready = true;
    return VR_OK;
}

// Render this rank's rows as up to BANDS bands of whole row tiles, each marched on its own stream and followed
// there by the device->host copy of its rows, so that only the last band's copy is exposed (33 MB over PCIe
// cost 0.7 ms per 1080p frame when copied after the whole march).  `full_frame` = 1: unpartitioned frame, the
// host buffer receives every row (vr_render); 0: the rank's row tiles land in their rows of a full host frame
// that other ranks fill too (vr_render_owned_to_host).  The device image is compact (owned rows only).
static int ensure_slot(vr_context::FrameSlot& f)
{
    if (f.done) return VR_OK;
    VR_CUDA(cudaEventCreate(&f.ev0));
    VR_CUDA(cudaEventCreate(&f.ev1));
    VR_CUDA(cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
    return VR_OK;
}

// Enqueues the frame -- march of every band, each followed in its own stream by the device->host copy of its rows --
// and returns without waiting.  `f.done` fires when the whole frame is in host memory, `f.ev0 -> f.ev1` brackets the
// marches.  Band b of the NEXT frame is ordered, by its stream, after band b's copy of this one: two frames can be in
// flight on the one device image, and the host's work for a frame overlaps the GPU's work on the one before.
static int render_banded_submit(vr_context* c, float* host_rgba, vr_context::FrameSlot& f)
{
    const int B = band_count((size_t)c->W * (size_t)compact_rows_of(c->H, c->rank, c->world, c->tile_rows));
    int rc = ensure_bands(c);
    if (rc != VR_OK) return rc;
    LaunchPlan plan;
    rc = make_plan(c, /*compact=*/1, &plan);
    if (rc != VR_OK) return rc;
    const int T = c->tile_rows;
    const int local_tiles = plan.local_rows / T;
    const int tiles_per_band = std::max(1, (local_tiles + B - 1) / B);
    const size_t row_floats = (size_t)c->W * 4;
    uint32_t launches = 0;
    int nb = 0;
    cudaError_t e = cudaEventRecord(f.ev0, c->band_stream[0]);             // band 0 is the first to start
    for (int b = 0; b < B && rc == VR_OK && e == cudaSuccess; ++b) {
        const int t0 = b * tiles_per_band, t1 = std::min(local_tiles, t0 + tiles_per_band);
        if (t0 >= t1) break;
        cudaStream_t bs = c->band_stream[b];
        rc = launch_march(c, plan, c->d_frame, t0 * T, t1 * T, bs, nullptr, nullptr);
        ++launches;
        if (e == cudaSuccess) e = cudaEventRecord(c->band_kdone[b], bs);
        // The band's tiles are T*W*16-byte chunks, contiguous in the compact device image and world*T rows apart in
        // the host frame: ONE strided 2-D copy per band (plus one plain copy if the frame ends in a partial tile).
        {
            const size_t tile_bytes = (size_t)T * row_floats * sizeof(float);
            int full = 0;
            for (int lt = t0; lt < t1; ++lt) if ((lt * c->world + c->rank) * T + T <= c->H) ++full;
            if (full > 0 && e == cudaSuccess)
                e = cudaMemcpy2DAsync(host_rgba + (size_t)(t0 * c->world + c->rank) * T * row_floats, tile_bytes * (size_t)c->world,
                                      c->d_frame + (size_t)t0 * T * row_floats, tile_bytes, tile_bytes, (size_t)full,
                                      cudaMemcpyDeviceToHost, bs);
            for (int lt = t0 + full; lt < t1 && e == cudaSuccess; ++lt) {
                const int y0 = (lt * c->world + c->rank) * T, rows = std::min(T, c->H - y0);
                if (rows > 0)
                    e = cudaMemcpyAsync(host_rgba + (size_t)y0 * row_floats, c->d_frame + (size_t)lt * T * row_floats,
                                        (size_t)rows * row_floats * sizeof(float), cudaMemcpyDeviceToHost, bs);
            }
        }
        if (e == cudaSuccess) e = cudaEventRecord(c->band_cdone[b], bs);
        nb = b + 1;
    }
    // join on the context's stream: it only ever carries these waits and records, so frames do not serialise on it
    for (int b = 0; b < nb && e == cudaSuccess; ++b) e = cudaStreamWaitEvent(c->stream, c->band_kdone[b], 0);
    if (e == cudaSuccess) e = cudaEventRecord(f.ev1, c->stream);            // every band's march has finished
    for (int b = 0; b < nb && e == cudaSuccess; ++b) e = cudaStreamWaitEvent(c->stream, c->band_cdone[b], 0);
    if (e == cudaSuccess) e = cudaEventRecord(f.done, c->stream);           // the frame is in host memory
    if (e != cudaSuccess || rc != VR_OK) {
        for (int b = 0; b < nb; ++b) cudaStreamSynchronize(c->band_stream[b]);
        cudaStreamSynchronize(c->stream);
        return e != cudaSuccess ? cuda_fail(e, "banded render") : rc;
    }
    f.launches = launches; f.kernel = plan.kernel; f.skip = plan.skip;
    return VR_OK;
}

static int frame_slot_wait(vr_context* c, vr_context::FrameSlot& f, vr_render_stats* stats)
{
    (void)c;
    VR_CUDA(cudaEventSynchronize(f.done));
    float ms = 0.f;
    VR_CUDA(cudaEventElapsedTime(&ms, f.ev0, f.ev1));
    if (stats) { stats->kernel_ms = ms; stats->kernel_launches = f.launches; stats->kernel_used = (uint32_t)f.kernel; stats->skip_used = f.skip ? 1u : 0u; }
    return VR_OK;
}

static int render_banded(vr_context* c, float* host_rgba, vr_render_stats* stats)
{
    // synchronous form: no frame may be in flight (the caller of vr_render_submit waits for its tickets first)
    vr_context::FrameSlot& f = c->slot[0];
    int rc = ensure_slot(f);
    if (rc != VR_OK) return rc;
    rc = render_banded_submit(c, host_rgba, f);
    if (rc != VR_OK) return rc;
    return frame_slot_wait(c, f, stats);
}

static bool banding_pays(const vr_context* c)
{
    static const bool no_bands = std::getenv("VR_NO_BANDS") != nullptr;
    // small frames copy in microseconds: four launches on four streams would cost more than they hide
    const size_t owned_px = (size_t)c->W * (size_t)compact_rows_of(c->H, c->rank, c->world, c->tile_rows);
    return !no_bands && owned_px >= ((size_t)1 << 17) && compact_rows_of(c->H, c->rank, c->world, c->tile_rows) >= 2 * c->tile_rows;
}

int vr_render(vr_context* c, float* host_rgba, vr_render_stats* stats)
{
    if (!c || !host_rgba) return fail(VR_ERR_INVALID, "vr_render: null argument");
    if (c->slot[0].pending || c->slot[1].pending) return fail(VR_ERR_INVALID, "vr_render: frames submitted with vr_render_submit are still in flight");
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    const size_t bytes = (size_t)c->W * c->H * 4 * sizeof(float);
    int rc;
    if (c->world == 1 && banding_pays(c)) {
        // the banded path leaves the frame in the context's buffer too (compact == full when world == 1)
        rc = render_banded(c, host_rgba, stats);
        if (rc != VR_OK) return rc;
    } else {
        if (c->world > 1) VR_CUDA(cudaMemsetAsync(c->d_frame, 0, bytes, c->stream));
        rc = render_common(c, c->d_frame, 0, c->stream, stats);
        if (rc != VR_OK) return rc;
        VR_CUDA(cudaMemcpyAsync(host_rgba, c->d_frame, bytes, cudaMemcpyDeviceToHost, c->stream));
        VR_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

// Multi-GPU end to end: this rank's row tiles straight into a FULL host frame (one buffer shared by all
// ranks, e.g. POSIX shared memory registered with cudaHostRegister in every process): N PCIe links carry
// the frame instead of rank 0's one, and nothing crosses NVLink.  Rows this rank does not own are not touched.
// Banded like vr_render: a band's tiles are copied while the following bands are still marching.
int vr_render_owned_to_host(vr_context* c, float* host_full_frame, vr_render_stats* stats)
{
    if (!c || !host_full_frame) return fail(VR_ERR_INVALID, "vr_render_owned_to_host: null argument");
    if (c->slot[0].pending || c->slot[1].pending) return fail(VR_ERR_INVALID, "vr_render_owned_to_host: frames submitted with vr_render_submit are still in flight");
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    int rc;
    if (banding_pays(c)) {
        rc = render_banded(c, host_full_frame, stats);
        if (rc != VR_OK) return rc;
    } else {
        rc = render_common(c, c->d_frame, /*compact=*/1, c->stream, stats);     // compact tiles fit in the frame buffer
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
    }
    if (stats) stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

// Pipelined form of vr_render / vr_render_owned_to_host: vr_render_submit enqueues the frame and returns a ticket,
// vr_render_wait blocks until that frame is complete in host memory.  Up to two frames may be in flight; the host's
// per-frame work (planning, launches, the consumer's hand-shake) then overlaps the GPU's work on the previous frame.
int vr_render_submit(vr_context* c, float* host_frame, uint32_t* ticket)
{
    if (!c || !host_frame || !ticket) return fail(VR_ERR_INVALID, "vr_render_submit: null argument");
    VR_CUDA(cudaSetDevice(c->device));
    const uint32_t t = c->next_ticket;
    vr_context::FrameSlot& f = c->slot[t % vr_context::SLOTS];
    if (f.pending) return fail(VR_ERR_INVALID, "vr_render_submit: two frames are already in flight; wait for the older ticket first");
    int rc = ensure_slot(f);
    if (rc != VR_OK) return rc;
    if (banding_pays(c)) {
        rc = render_banded_submit(c, host_frame, f);
        if (rc != VR_OK) return rc;
    } else {
        // small frames: one launch and the copies of the owned tiles, all in the context's stream
        LaunchPlan plan;
        rc = make_plan(c, /*compact=*/1, &plan);
        if (rc != VR_OK) return rc;
        VR_CUDA(cudaEventRecord(f.ev0, c->stream));
        rc = launch_march(c, plan, c->d_frame, 0, plan.local_rows, c->stream, nullptr, nullptr);
        if (rc != VR_OK) return rc;
        VR_CUDA(cudaEventRecord(f.ev1, c->stream));
        const int tiles = (c->H + c->tile_rows - 1) / c->tile_rows;
        const size_t row_floats = (size_t)c->W * 4;
        int local_tile = 0;
        for (int tl = c->rank; tl < tiles; tl += c->world, ++local_tile) {
            const int y0 = tl * c->tile_rows, rows = std::min(c->tile_rows, c->H - y0);
            VR_CUDA(cudaMemcpyAsync(host_frame + (size_t)y0 * row_floats, c->d_frame + (size_t)local_tile * c->tile_rows * row_floats,
                                    (size_t)rows * row_floats * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        }
        VR_CUDA(cudaEventRecord(f.done, c->stream));
        f.launches = plan.local_rows > 0 ? 1u : 0u; f.kernel = plan.kernel; f.skip = plan.skip;
    }
    f.pending = true;
    c->slot_ticket[t % vr_context::SLOTS] = t;
    c->next_ticket = t + 1;
    *ticket = t;
    return VR_OK;
}

int vr_render_wait(vr_context* c, uint32_t ticket, vr_render_stats* stats)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_render_wait: null context");
    vr_context::FrameSlot& f = c->slot[ticket % vr_context::SLOTS];
    if (!f.pending || c->slot_ticket[ticket % vr_context::SLOTS] != ticket) return fail(VR_ERR_INVALID, "vr_render_wait: no such frame in flight");
    VR_CUDA(cudaSetDevice(c->device));
    const int rc = frame_slot_wait(c, f, stats);
    f.pending = false;
    return rc;
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

// a barrier wait of an earlier frame gave up: refuse to go on until the caller has looked (vr_peer_frame_reset)
static int peer_error_check(const vr_context* c, const char* who)
{
    const unsigned int e = *(volatile unsigned int*)c->h_peer_error;
    if (e == 0) return VR_OK;
    char buf[256];
    std::snprintf(buf, sizeof buf, "%s: an earlier peer-frame wait (%s) timed out -- the frame may be torn; call vr_peer_frame_reset",
                  who, e == 1 ? "arrivals" : "release");
    return fail(VR_ERR_TIMEOUT, buf);
}

int vr_peer_frame_arrive(vr_context* c, float* d_target_frame, uint32_t frame_no, int world, int is_owner, void* cuda_stream)
{
    if (!c || !d_target_frame || world < 1) return fail(VR_ERR_INVALID, "vr_peer_frame_arrive: bad argument");
    int rc = peer_error_check(c, "vr_peer_frame_arrive");
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream;
    unsigned int* sync = frame_sync_words(c, d_target_frame);
    vr::peer_signal_kernel<<<1, 1, 0, s>>>(sync);
    if (is_owner) vr::peer_wait_kernel<<<1, 1, 0, s>>>(sync, 0, frame_no * (uint32_t)world, peer_timeout_ns(), c->d_peer_error);
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

int vr_peer_frame_release(vr_context* c, float* d_target_frame, uint32_t frame_no, int is_owner, void* cuda_stream)
{
    if (!c || !d_target_frame) return fail(VR_ERR_INVALID, "vr_peer_frame_release: bad argument");
    int rc = peer_error_check(c, "vr_peer_frame_release");
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream;
    unsigned int* sync = frame_sync_words(c, d_target_frame);
    if (is_owner) vr::peer_release_kernel<<<1, 1, 0, s>>>(sync, frame_no);
    else          vr::peer_wait_kernel<<<1, 1, 0, s>>>(sync, 1, frame_no, peer_timeout_ns(), c->d_peer_error);
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

// render + arrive: the march kernel's last CTA publishes this rank's arrival (no signal kernel); kernels without
// that epilogue (the generic loop) and ranks that own no row fall back to the one-thread signal kernel
int vr_render_peer(vr_context* c, float* d_target_frame, uint32_t frame_no, int world, int is_owner, void* cuda_stream,
                   vr_render_stats* stats)
{
    if (!c || !d_target_frame || world < 1) return fail(VR_ERR_INVALID, "vr_render_peer: bad argument");
    int rc = peer_error_check(c, "vr_render_peer");
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(c->device));
    const auto t0 = std::chrono::steady_clock::now();
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream;
    unsigned int* sync = frame_sync_words(c, d_target_frame);
    LaunchPlan plan;
    rc = make_plan(c, /*compact=*/0, &plan);
    if (rc != VR_OK) return rc;
    bool signalled = false;
    c->done_slot = (int)(frame_no & 3u);          // concurrent frames (different streams) must differ in frame_no mod 4
    // stats == NULL: asynchronous -- nothing is waited for; the march's bracket goes into a ring of event pairs that
    // vr_peer_kernel_ms reads later (frames then run back to back, the host never idles the GPU between them)
    cudaEvent_t e0 = c->ev0, e1 = c->ev1;
    if (!stats) {
        const int k = (int)(frame_no % vr_context::PEER_EVENTS);
        if (!c->peer_ev[k][0]) { VR_CUDA(cudaEventCreate(&c->peer_ev[k][0])); VR_CUDA(cudaEventCreate(&c->peer_ev[k][1])); }
        e0 = c->peer_ev[k][0]; e1 = c->peer_ev[k][1]; c->peer_ev_frame[k] = frame_no;
    }
    VR_CUDA(cudaEventRecord(e0, s));
    rc = launch_march(c, plan, d_target_frame, 0, plan.local_rows, s, &sync[0], &signalled);
    if (rc != VR_OK) return rc;
    uint32_t launches = plan.local_rows > 0 ? 1u : 0u;
    if (!signalled) { vr::peer_signal_kernel<<<1, 1, 0, s>>>(sync); ++launches; }
    VR_CUDA(cudaEventRecord(e1, s));
    if (is_owner) { vr::peer_wait_kernel<<<1, 1, 0, s>>>(sync, 0, frame_no * (uint32_t)world, peer_timeout_ns(), c->d_peer_error); ++launches; }
    VR_CUDA(cudaGetLastError());
    if (!stats) return VR_OK;
    VR_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    VR_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    fill_stats(stats, ms, launches, plan);
    stats->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VR_OK;
}

// march-kernel time of an asynchronous vr_render_peer call (stats == NULL) for `frame_no`; waits for that kernel
int vr_peer_kernel_ms(vr_context* c, uint32_t frame_no, float* ms)
{
    if (!c || !ms) return fail(VR_ERR_INVALID, "vr_peer_kernel_ms: null argument");
    const int k = (int)(frame_no % vr_context::PEER_EVENTS);
    if (!c->peer_ev[k][0] || c->peer_ev_frame[k] != frame_no) return fail(VR_ERR_INVALID, "vr_peer_kernel_ms: that frame's bracket is no longer kept (ring of 64)");
    VR_CUDA(cudaSetDevice(c->device));
    VR_CUDA(cudaEventSynchronize(c->peer_ev[k][1]));
    VR_CUDA(cudaEventElapsedTime(ms, c->peer_ev[k][0], c->peer_ev[k][1]));
    return VR_OK;
}

// the frame owner's side of the barrier on a stream of its own choice (e.g. a consumer stream, so that the owner's
// march kernels are not held back by the slowest rank): wait until `frame_no * world` arrivals have been published
int vr_peer_frame_wait_arrivals(vr_context* c, float* d_target_frame, uint32_t frame_no, int world, void* cuda_stream)
{
    if (!c || !d_target_frame || world < 1) return fail(VR_ERR_INVALID, "vr_peer_frame_wait_arrivals: bad argument");
    int rc = peer_error_check(c, "vr_peer_frame_wait_arrivals");
    if (rc != VR_OK) return rc;
    VR_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->stream;
    vr::peer_wait_kernel<<<1, 1, 0, s>>>(frame_sync_words(c, d_target_frame), 0, frame_no * (uint32_t)world, peer_timeout_ns(), c->d_peer_error);
    VR_CUDA(cudaGetLastError());
    return VR_OK;
}

int vr_peer_frame_reset(vr_context* c, float* d_target_frame)
{
    if (!c) return fail(VR_ERR_INVALID, "vr_peer_frame_reset: null context");
    VR_CUDA(cudaSetDevice(c->device));
    VR_CUDA(cudaDeviceSynchronize());
    *c->h_peer_error = 0;
    if (d_target_frame) VR_CUDA(cudaMemset(frame_sync_words(c, d_target_frame) + 2, 0, sizeof(unsigned int)));
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
    const dim3 grid((c->W + vr::DIRECT_BLOCK_W - 1) / vr::DIRECT_BLOCK_W, (plan.local_rows + vr::DIRECT_BLOCK_H - 1) / vr::DIRECT_BLOCK_H);
    plan.fc.tc_div_mode = vr::DIV_IEEE;
    rc = plan.local_rows == 0 ? VR_OK
       : c->bpv == 1 ? launch_direct_t<uint8_t, true>(c, plan, args, grid, c->stream)
                     : launch_direct_t<uint16_t, true>(c, plan, args, grid, c->stream);
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
