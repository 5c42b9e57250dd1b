// marchlab.cu -- native kernel lab: runs march-kernel variants on BASELINE's headline workload
// (1024^3 uint16 synthetic mix, 1920x1080, camera K2, alpha 0.02, trilinear) on one GPU, checks
// each variant bit-for-bit against the baseline direct kernel (which the pytest suite checks
// against the CPU oracle) and prints CUDA-event timings.  Development tool, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -I../../include -I../../volume-renderer_b200/csrc -I. marchlab.cu -o marchlab
// Holds the round-1 development kernels that were measured slower than the product's two (kernel_lab.cuh: LSU "fast" /
// "packed", two-gather, non-pipelined and two-ray z-pair, hybrid TEX+LSU; kernel_windowed.cuh: TMA-staged windows).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#include "volren_b200.h"
#include "frame.h"
#include "march_device.cuh"
#include "kernel_direct.cuh"
#include "kernel_lab.cuh"
#include "kernels_aux.cuh"
#include "lab_aux.cuh"
#include "kernel_windowed.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

using namespace vr;

static const float K0[21] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,3,1, 0,0,3,1, 3.7320508f};
static const float K1[21] = {0.81915206f,0,-0.5735764f,0, -0.28678817f,0.8660254f,-0.409576f,0, 0.49673176f,0.49999994f,0.7094065f,0,
                             1.4901954f,1.4999999f,2.1282196f,1, 1.4901954f,1.4999999f,2.1282196f,1, 3.7320508f};
static const float K2[21] = {-0.9396927f,0,0.34202015f,0, 0.11697778f,0.9396926f,0.32139382f,0, -0.3213938f,0.34202012f,-0.8830222f,0,
                             -0.51423013f,0.5472323f,-1.4128356f,1, -0.51423013f,0.5472323f,-1.4128356f,1, 3.7320508f};

struct Lab {
    int W = 1920, H = 1080, N = 1024;
    int rows = 1080;          // local rows rendered (emulated rank of a partition)
    uint16_t* d_pad = nullptr; uint32_t pitch = 0; uint64_t slice = 0;
    uint32_t* d_pairs = nullptr;
    cudaArray_t arr = nullptr; cudaTextureObject_t tex = 0;
    cudaArray_t arrf = nullptr; cudaTextureObject_t texf = 0;
    float *d_ref = nullptr, *d_out = nullptr;
    FrameConsts fc;
    cudaEvent_t e0, e1;
};

template <typename F>
static float time_kernel(Lab& L, F launch, int reps = 5)
{
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(L.e0));
        launch();
        CK(cudaEventRecord(L.e1));
        CK(cudaEventSynchronize(L.e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, L.e0, L.e1));
        if (ms < best) best = ms;
    }
    return best;
}

static void check(Lab& L, const char* name, float ms, double gsamples)
{
    const size_t n = (size_t)L.W * L.H * 4;
    static std::vector<uint32_t> a, b;
    a.resize(n); b.resize(n);
    CK(cudaMemcpy(a.data(), L.d_ref, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), L.d_out, n * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad += a[i] != b[i];
    printf("%-44s %8.3f ms  %7.1f Mrays/s  %6.1f Gsamples/s  %s (%zu values differ)\n", name, ms,
           L.W * (double)L.H / ms / 1e3, gsamples / ms * 1e3 / 1e9 * 1e0, bad ? "MISMATCH" : "bit-exact", bad);
    CK(cudaMemset(L.d_out, 0xff, n * 4));
}

static const char* g_only = nullptr;   // run only variants whose name contains this

static int g_persist = 0;               // >0: persistent grid with this many CTAs per SM
static int g_sms = 148;
static int g_center = 0;
static unsigned int* g_counter = nullptr;

template <typename T, int FILTER, int TCDIV, int WIN, int FM, int RAYS, bool PAIRS = false, bool PIPE = false>
static void run_fast(Lab& L, const char* name, double samples)
{
    if (g_only && !strstr(name, g_only)) return;
    FastArgs a{};
    a.vol = PAIRS ? (const void*)L.d_pairs : (const void*)L.d_pad; a.pitch = L.pitch; a.slice_lo = (uint32_t)L.slice; a.out = L.d_out; a.local_rows = L.rows;
    const int cols = (L.W + RAYS - 1) / RAYS;
    a.local_rows = L.rows;
    dim3 grid((cols + 31) / 32, (L.rows + 7) / 8), block(FAST_THREADS);
    const float ms = time_kernel(L, [&] { march_fast_kernel<T, FILTER, TCDIV, WIN, FM, RAYS, PAIRS, PIPE><<<grid, block>>>(L.fc, a); });
    check(L, name, ms, samples);
}

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP>
static void run_packed(Lab& L, const char* name, double samples)
{
    if (g_only && !strstr(name, g_only)) return;
    FastArgs a{};
    a.vol = L.d_pad; a.pitch = L.pitch; a.slice_lo = (uint32_t)L.slice; a.out = L.d_out; a.local_rows = L.rows;
    dim3 grid((L.W + 31) / 32, (L.rows + 7) / 8), block(256);
    const float ms = time_kernel(L, [&] { march_packed_kernel<T, TCDIV, WIN, UNIT, NOCAP><<<grid, block>>>(L.fc, a); });
    check(L, name, ms, samples);
}

__global__ void u16_to_f32_kernel(const uint16_t* __restrict__ s, float* __restrict__ d, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = (float)s[i];
}

template <typename T, int TCDIV, int WIN, bool UNIT, bool NOCAP, bool FLOATTEX = false>
static void run_texgather(Lab& L, const char* name, double samples)
{
    if (g_only && !strstr(name, g_only)) return;
    TexArgs a{};
    a.tex = FLOATTEX ? L.texf : L.tex; a.out = L.d_out; a.local_rows = L.rows;
    dim3 grid((L.W + 31) / 32, (L.rows + 7) / 8), block(256);
    const float ms = time_kernel(L, [&] { march_texgather_kernel<T, TCDIV, WIN, UNIT, NOCAP, FLOATTEX><<<grid, block>>>(L.fc, a); });
    check(L, name, ms, samples);
}

template <typename T>
static void run_windowed(Lab& L, const char* name, double samples, int tcdiv, int win, int ctas_per_sm, int sms, bool stats = false)
{
    if (g_only && !strstr(name, g_only)) return;
    static WindowedState st;
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
        CK(cudaEventRecord(L.e0));
        if (launch_windowed_t<T>(st, L.fc, L.d_pad, L.pitch, L.slice, L.N, L.N, L.d_out, L.rows, sms, tcdiv, win, 0, stats, ctas_per_sm) != 0) {
            printf("%s: launch failed: %s\n", name, st.err); return;
        }
        CK(cudaEventRecord(L.e1));
        cudaError_t e = cudaEventSynchronize(L.e1);
        if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e)); exit(2); }
        float ms; CK(cudaEventElapsedTime(&ms, L.e0, L.e1));
        if (ms < best) best = ms;
    }
    unsigned long long stt[3] = {0, 0, 0};
    if (stats) CK(cudaMemcpy(stt, st.d_stats, sizeof stt, cudaMemcpyDeviceToHost));
    if (stats) printf("    [windowed stats: smem samples %llu, global-fallback samples %llu (%.2f%%), windows %llu]\n", stt[0], stt[1],
           100.0 * stt[1] / (double)(stt[0] + stt[1] + 1), stt[2]);
    check(L, name, best, samples);
}

int main(int argc, char** argv)
{
    Lab L;
    const char* camname = argc > 1 ? argv[1] : "K2";
    const float alpha = argc > 2 ? (float)atof(argv[2]) : 0.02f;
    const int filter = argc > 3 ? atoi(argv[3]) : 1;
    if (argc > 4) L.N = atoi(argv[4]);
    if (argc > 5) g_only = argv[5];
    const int world = argc > 6 ? atoi(argv[6]) : 1;
    if (argc > 7) g_persist = atoi(argv[7]);
    if (argc > 8) g_center = atoi(argv[8]);
    const float* cam = !strcmp(camname, "K0") ? K0 : (!strcmp(camname, "K1") ? K1 : K2);
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs; workload %d^3 u16, %dx%d, camera %s, alpha %g, filter %d\n", prop.name, prop.multiProcessorCount,
           L.N, L.W, L.H, camname, alpha, filter);
    CK(cudaEventCreate(&L.e0)); CK(cudaEventCreate(&L.e1));
    const int N = L.N;
    const uint64_t nvox = (uint64_t)N * N * N;
    uint16_t* d_src; CK(cudaMalloc(&d_src, nvox * 2));
    synth_mix_kernel<uint16_t><<<prop.multiProcessorCount * 16, 256>>>(d_src, N, N, N, 4095, 0x5EED0004u, 1);
    L.pitch = (uint32_t)(((uint64_t)(N + 2) * 2 + 15) / 16 * 16 / 2);
    L.slice = (uint64_t)L.pitch * (N + 2);
    CK(cudaMalloc(&L.d_pad, L.slice * (N + 2) * 2 + 256));
    lab::pad_volume_kernel<uint16_t><<<prop.multiProcessorCount * 16, 256>>>(d_src, L.d_pad, N, N, N, L.pitch);
    {   // layered 2-D array + point-sampling texture object for the gather variant
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
        CK(cudaMalloc3DArray(&L.arr, &cd, make_cudaExtent(N, N, N), cudaArrayLayered));
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(d_src, (size_t)N * 2, N, N);
        cp.dstArray = L.arr; cp.extent = make_cudaExtent(N, N, N); cp.kind = cudaMemcpyDeviceToDevice;
        CK(cudaMemcpy3D(&cp));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = L.arr;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        CK(cudaCreateTextureObject(&L.tex, &rd, &td, nullptr));
    }
    {   // float-texel layered array (lab experiment)
        float* d_f; CK(cudaMalloc(&d_f, nvox * 4));
        u16_to_f32_kernel<<<prop.multiProcessorCount * 16, 256>>>(d_src, d_f, nvox);
        cudaChannelFormatDesc cd = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
        CK(cudaMalloc3DArray(&L.arrf, &cd, make_cudaExtent(N, N, N), cudaArrayLayered));
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(d_f, (size_t)N * 4, N, N);
        cp.dstArray = L.arrf; cp.extent = make_cudaExtent(N, N, N); cp.kind = cudaMemcpyDeviceToDevice;
        CK(cudaMemcpy3D(&cp));
        CK(cudaFree(d_f));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = L.arrf;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        CK(cudaCreateTextureObject(&L.texf, &rd, &td, nullptr));
    }
    CK(cudaMalloc(&L.d_pairs, L.slice * (N + 2) * 4 + 256));
    lab::pad_pairs_kernel<uint16_t, uint32_t><<<prop.multiProcessorCount * 16, 256>>>(d_src, L.d_pairs, N, N, N, L.pitch);
    CK(cudaDeviceSynchronize());
    CK(cudaFree(d_src));
    CK(cudaMalloc(&L.d_ref, (size_t)L.W * L.H * 16)); CK(cudaMalloc(&L.d_out, (size_t)L.W * L.H * 16));

    vr_params p; memset(&p, 0, sizeof p);
    p.alpha_scale = alpha; p.min_val = 0; p.max_val = 4095; p.filter = filter; p.step_scale = 1.0f;
    const int32_t dim[3] = {N, N, N}; const float vs[3] = {1, 1, 1};
    memset(&L.fc, 0, sizeof L.fc);
    compute_frame_consts(L.fc, L.W, L.H, dim, vs, cam, p);
    L.fc.rank = 0; L.fc.world = world; L.fc.tile_rows = 16; L.fc.compact = 0;
    L.rows = ((L.H + 15) / 16 + world - 1) / world * 16;
    g_sms = prop.multiProcessorCount;
    printf("emulating rank 0 of %d: %d local rows; persistent CTAs/SM: %d\n", world, L.rows, g_persist);
    printf("tc_div_mode %d (0 = exact reciprocal), denom %g %g %g, step %g\n", L.fc.tc_div_mode, L.fc.denom[0], L.fc.denom[1], L.fc.denom[2], L.fc.step);

    // sample count through the instrumented kernel (also the reference image source)
    unsigned long long* d_cnt; CK(cudaMalloc(&d_cnt, 24)); CK(cudaMemset(d_cnt, 0, 24));
    unsigned int* d_bits; CK(cudaMalloc(&d_bits, (nvox + 31) / 32 * 4)); CK(cudaMemset(d_bits, 0, (nvox + 31) / 32 * 4));
    DirectArgs da{}; da.vol = L.d_pad; da.pitch = L.pitch; da.slice = L.slice; da.out = L.d_ref; da.local_rows = L.rows;
    da.touch_bits = d_bits; da.counters = d_cnt;
    dim3 dgrid((L.W + 31) / 32, (L.rows + 7) / 8), dblock(256);
    march_direct_kernel<uint16_t, 0, DIV_IEEE, true, true><<<dgrid, dblock>>>(L.fc, da);
    CK(cudaDeviceSynchronize());
    unsigned long long cnt[3]; CK(cudaMemcpy(cnt, d_cnt, 24, cudaMemcpyDeviceToHost));
    const double samples = (double)cnt[0];
    printf("samples per frame %.0f, rays hit %llu\n", samples, cnt[1]);
    CK(cudaFree(d_bits));

    da.touch_bits = nullptr; da.counters = nullptr;
    float ms;
    if (filter == 1) {
        ms = time_kernel(L, [&] { march_direct_kernel<uint16_t, 1, DIV_RECIP_EXACT, false, false><<<dgrid, dblock>>>(L.fc, da); });
    } else {
        ms = time_kernel(L, [&] { march_direct_kernel<uint16_t, 0, DIV_RECIP_EXACT, false, false><<<dgrid, dblock>>>(L.fc, da); });
    }
    CK(cudaMemcpy(L.d_out, L.d_ref, (size_t)L.W * L.H * 16, cudaMemcpyDeviceToDevice));
    check(L, "direct baseline (round-1 first kernel)", ms, samples);

    if (filter == 1) {
        run_windowed<uint16_t>(L, "windowed TMA 1ray covers0 (auto, stats)", samples, DIV_RECIP_EXACT, WIN_COVERS0, 0, prop.multiProcessorCount, true);
        run_windowed<uint16_t>(L, "windowed TMA 1ray covers0 (auto occupancy)", samples, DIV_RECIP_EXACT, WIN_COVERS0, 0, prop.multiProcessorCount);
        run_windowed<uint16_t>(L, "windowed TMA 1ray covers0 (2 CTA/SM)", samples, DIV_RECIP_EXACT, WIN_COVERS0, 2, prop.multiProcessorCount);
        run_windowed<uint16_t>(L, "windowed TMA 1ray covers0 (4 CTA/SM)", samples, DIV_RECIP_EXACT, WIN_COVERS0, 4, prop.multiProcessorCount);
        run_texgather<uint16_t, DIV_RECIP_EXACT, WIN_COVERS0, true, true>(L, "texgather packed covers0 UNIT NOCAP", samples);
        run_texgather<uint16_t, DIV_RECIP_EXACT, WIN_COVERS0, true, true, true>(L, "texgather FLOAT texels covers0 UNIT NOCAP", samples);
        run_texgather<uint16_t, DIV_RECIP_EXACT, WIN_CLAMP, false, false>(L, "texgather packed clamp", samples);
        run_packed<uint16_t, DIV_RECIP_EXACT, WIN_COVERS0, false, false>(L, "packed-in-ray covers0", samples);
        run_packed<uint16_t, DIV_RECIP_EXACT, WIN_COVERS0, true, false>(L, "packed-in-ray covers0 UNIT", samples);
        run_packed<uint16_t, DIV_RECIP_EXACT, WIN_COVERS0, true, true>(L, "packed-in-ray covers0 UNIT NOCAP", samples);
        run_packed<uint16_t, DIV_RECIP_EXACT, WIN_CLAMP, false, false>(L, "packed-in-ray clamp", samples);
        run_packed<uint16_t, DIV_MARKSTEIN, WIN_CLAMP, false, false>(L, "packed-in-ray clamp markstein-tc", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 1, false, true>(L, "pipe u16   1ray  covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 2, false, true>(L, "pipe u16   2rays covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 1, true, true>(L, "pipe pairs 1ray  covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 2, true, true>(L, "pipe pairs 2rays covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_CLAMP, FLOOR_XU1, 1, false, true>(L, "pipe u16   1ray  clamp   floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 1, true>(L, "pairs 1ray  covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 2, true>(L, "pairs 2rays covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU2, 2, true>(L, "pairs 2rays covers0 floor=XU2", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU2, 1>(L, "fast 1ray  covers0 floor=XU2", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 1>(L, "fast 1ray  covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_MAGIC, 1>(L, "fast 1ray  covers0 floor=MAGIC", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU2, 2>(L, "fast 2rays covers0 floor=XU2", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU1, 2>(L, "fast 2rays covers0 floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_MAGIC, 2>(L, "fast 2rays covers0 floor=MAGIC", samples);
        run_fast<uint16_t, 1, DIV_RECIP_EXACT, WIN_CLAMP, FLOOR_XU1, 2>(L, "fast 2rays clamp   floor=XU1", samples);
        run_fast<uint16_t, 1, DIV_MARKSTEIN, WIN_CLAMP, FLOOR_XU1, 2>(L, "fast 2rays clamp markstein-tc floor=XU1", samples);
    } else {
        run_fast<uint16_t, 0, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU2, 1>(L, "fast 1ray  nearest covers0", samples);
        run_fast<uint16_t, 0, DIV_RECIP_EXACT, WIN_COVERS0, FLOOR_XU2, 2>(L, "fast 2rays nearest covers0", samples);
        run_fast<uint16_t, 0, DIV_RECIP_EXACT, WIN_CLAMP, FLOOR_XU2, 2>(L, "fast 2rays nearest clamp", samples);
    }
    return 0;
}
