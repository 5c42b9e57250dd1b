// issue_rates.cu -- measures warp-instruction issue rates (per SM per clock) of the instruction
// classes the ray-march inner loop is made of, on B200.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

#define ITERS 2048
#define NACC 8

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, const float* in, long long* cycles)
{
    const int t = threadIdx.x;
    float a[NACC], b = in[t & 31], c = in[32 + (t & 31)];
    u64 p[NACC], pb = pack(b, c), pc = pack(c, b);
    int ia[NACC];
    __shared__ unsigned short sm[8192];
    for (int i = t; i < 8192; i += blockDim.x) sm[i] = (unsigned short)(i * 7);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = in[64 + i] + t; p[i] = pack(a[i], a[i] + 1.f); ia[i] = t * 3 + i; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) { a[i] = __fmaf_rn(a[i], b, c); a[i] = __fmaf_rn(a[i], c, b); }                 // 2 FFMA
            if (MODE == 1) { p[i] = fma2(p[i], pb, pc); p[i] = fma2(p[i], pc, pb); }                       // 2 FFMA2
            if (MODE == 2) { p[i] = add2(p[i], pb); p[i] = add2(p[i], pc); }                               // 2 FADD2
            if (MODE == 3) { a[i] = __fadd_rn(a[i], b); a[i] = __fadd_rn(a[i], c); }                       // 2 FADD
            if (MODE == 4) { a[i] = __fmaf_rn(a[i], b, c); ia[i] = (ia[i] + it) ^ i; }                      // FFMA + 2 ALU
            if (MODE == 5) { p[i] = fma2(p[i], pb, pc); ia[i] = (ia[i] + it) ^ i; }                         // FFMA2 + 2 ALU
            if (MODE == 6) { ia[i] = __float2int_rd(a[i]); a[i] = a[i] + 0.37f; }                           // F2I + FADD
            if (MODE == 7) { a[i] = floorf(a[i]) + 0.37f; }                                                 // FRND + FADD
            if (MODE == 8) { a[i] = (float)ia[i] * 1.0001f; ia[i] += it; }                                  // I2FP + FMUL + IADD
            if (MODE == 9) { a[i] = __uint_as_float(0x4B000000u | (ia[i] & 0xffff)) - 8388608.0f; ia[i] += it; }   // LOP3 + FADD + IADD
            if (MODE == 10) { ia[i] = sm[(ia[i] & 8191)] + it; }                                            // LDS.U16 dependent chain (gather)
            if (MODE == 11) { ia[i] += sm[((t * 2 + i * 64 + it) & 8191)]; }                                // LDS.U16 conflict-free-ish
            if (MODE == 12) { ia[i] += sm[((t * 37 + i * 531 + it * 3) & 8191)]; }                          // LDS.U16 scattered
            if (MODE == 13) { a[i] = fmaxf(fminf(a[i], c), b); a[i] += 0.1f; }                              // 2 FMNMX + FADD
        }
    }
    long long t1 = clock64();
    float s = 0; int si = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float x, y; unpack(p[i], x, y); s += a[i] + x + y; si += ia[i]; }
    out[blockIdx.x * blockDim.x + t] = s + si;
    if (t == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_inner, float* out, float* in, long long* cyc, int sms)
{
    const int blocks = sms * 4;
    bench<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, 256>>>(out, in, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    // warp-instructions per SM: 4 CTAs * 8 warps * ITERS * NACC * instr_per_inner
    const double winstr = 4.0 * 8 * ITERS * NACC * instr_per_inner;
    printf("%-28s cycles/CTA %8lld  warp-instr/clk/SM %.3f  (%.3f ms)\n", name, h[0], winstr / (double)h[0], ms);
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float *out, *in; long long* cyc;
    cudaMalloc(&out, sms * 4 * 256 * sizeof(float)); cudaMalloc(&in, 4096); cudaMalloc(&cyc, sms * 4 * sizeof(long long));
    float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0f + i * 1e-3f;
    cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
    printf("%s, %d SMs\n", prop.name, sms);
    run<0>("FFMA x2", 2, out, in, cyc, sms);
    run<1>("FFMA2 x2", 2, out, in, cyc, sms);
    run<2>("FADD2 x2", 2, out, in, cyc, sms);
    run<3>("FADD x2", 2, out, in, cyc, sms);
    run<4>("FFMA + 2 ALU", 3, out, in, cyc, sms);
    run<5>("FFMA2 + 2 ALU", 3, out, in, cyc, sms);
    run<6>("F2I.FLOOR + FADD", 2, out, in, cyc, sms);
    run<7>("FRND.FLOOR + FADD", 2, out, in, cyc, sms);
    run<8>("I2FP + FMUL + IADD", 3, out, in, cyc, sms);
    run<9>("LOP3 + FADD + IADD(+LOP)", 4, out, in, cyc, sms);
    run<10>("LDS.U16 dependent + 2 ALU", 3, out, in, cyc, sms);
    run<11>("LDS.U16 linear + 2 ALU", 3, out, in, cyc, sms);
    run<12>("LDS.U16 scattered + 2 ALU", 3, out, in, cyc, sms);
    run<13>("2 FMNMX + FADD", 3, out, in, cyc, sms);
    return 0;
}
