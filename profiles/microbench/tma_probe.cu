// tma_probe.cu -- minimal check of cp.async.bulk.tensor.3d tile loads of uint16 boxes on sm_100a
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BX, int BY>
__global__ void probe(const __grid_constant__ CUtensorMap tm, const CUtensorMap* gtm, int use_global, int x, int y, int z, uint16_t* out)
{
    __shared__ __align__(128) uint16_t buf[BX * BY];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* d = use_global ? gtm : &tm;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(BX * BY * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(s32(buf)), "l"(d), "r"(s32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) out[i] = buf[i];
}

template <int BX, int BY>
int run(PFN enc, uint16_t* d_vol, int px, int py, int pz, int use_global, int x, int y, int z, const std::vector<uint16_t>& h)
{
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)px, (cuuint64_t)py, (cuuint64_t)pz};
    const cuuint64_t strides[2] = {(cuuint64_t)px * 2, (cuuint64_t)px * py * 2};
    const cuuint32_t box[3] = {BX, BY, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d_vol, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%dx1 desc-in-%s at (%d,%d,%d): encode rc %d; ", BX, BY, use_global ? "global" : "param", x, y, z, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    CUtensorMap* d_tm; CK(cudaMalloc(&d_tm, sizeof tm)); CK(cudaMemcpy(d_tm, &tm, sizeof tm, cudaMemcpyHostToDevice));
    uint16_t* d_out; CK(cudaMalloc(&d_out, BX * BY * 2)); CK(cudaMemset(d_out, 0xee, BX * BY * 2));
    probe<BX, BY><<<1, 128>>>(tm, d_tm, use_global, x, y, z, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<uint16_t> o(BX * BY); CK(cudaMemcpy(o.data(), d_out, BX * BY * 2, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int j = 0; j < BY; ++j) for (int i = 0; i < BX; ++i) {
        const int gx = x + i, gy = y + j;
        const uint16_t exp = (gx >= 0 && gx < px && gy >= 0 && gy < py && z >= 0 && z < pz) ? h[((size_t)z * py + gy) * px + gx] : 0;
        bad += o[j * BX + i] != exp;
    }
    printf("kernel ok, %d mismatches\n", bad);
    return bad != 0;
}

int main(int argc, char** argv)
{
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    PFN enc = (PFN)p;
    const int px = 1032, py = 66, pz = 20;
    std::vector<uint16_t> h((size_t)px * py * pz);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i * 2654435761u >> 13);
    uint16_t* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    const int bx = argc > 1 ? atoi(argv[1]) : 32, by = argc > 2 ? atoi(argv[2]) : 16;
    const int x = argc > 3 ? atoi(argv[3]) : 8, y = argc > 4 ? atoi(argv[4]) : 4, z = argc > 5 ? atoi(argv[5]) : 3;
    const int g = argc > 6 ? atoi(argv[6]) : 0;
    if (bx == 32 && by == 16) return run<32, 16>(enc, d, px, py, pz, g, x, y, z, h);
    if (bx == 24 && by == 24) return run<24, 24>(enc, d, px, py, pz, g, x, y, z, h);
    if (bx == 24 && by == 16) return run<24, 16>(enc, d, px, py, pz, g, x, y, z, h);
    if (bx == 32 && by == 24) return run<32, 24>(enc, d, px, py, pz, g, x, y, z, h);
    if (bx == 16 && by == 16) return run<16, 16>(enc, d, px, py, pz, g, x, y, z, h);
    if (bx == 40 && by == 24) return run<40, 24>(enc, d, px, py, pz, g, x, y, z, h);
    if (bx == 8 && by == 8) return run<8, 8>(enc, d, px, py, pz, g, x, y, z, h);
    printf("unsupported box\n");
    return 1;
}
