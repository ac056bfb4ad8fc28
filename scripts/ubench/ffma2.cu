// Issue rate of the fp32 FMA forms on sm_100a: scalar FFMA (3 register operands), FFMA2 (fma.rn.f32x2) and HFMA2.BF16, with
// 8 independent accumulator chains per thread; prints FMAs per clock per SM for 4 / 8 / 16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o build/ffma2 scripts/ubench/ffma2.cu && build/ffma2
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[8], b = 1.0001f + threadIdx.x * 1e-6f, c = 0.5f;
    unsigned long long A[8], B, Cc;
    unsigned int H[8], HB = 0x3f803f80u, HC = 0x3f003f00u;
    for (int i = 0; i < 8; ++i) { a[i] = i + threadIdx.x; A[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); H[i] = 0x3f803f80u + i; }
    B = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    Cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fmaf(a[i], b, c);
                else if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(A[i]) : "l"(A[i]), "l"(B), "l"(Cc));
                else asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(H[i]) : "r"(H[i]), "r"(HB), "r"(HC));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)A[i]) + __uint_as_float(H[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[3] = {"FFMA (scalar)", "FFMA2 (f32x2)", "HFMA2.BF16"};
    for (int mode = 0; mode < 3; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
                else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
                else k<2><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double instr = (double)iters * 32 * warps;                       // warp instructions per SM
            const double fma_per_instr = mode == 0 ? 32 : 64;
            printf("%-14s %2d warps/SM: %8lld cycles, %.2f warp-instr/clk/SM, %.1f FMA/clk/SM\n", names[mode], warps, c,
                   instr / c, instr * fma_per_instr / c);
        }
    return 0;
}
