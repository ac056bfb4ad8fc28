// Issue rate of the special-function forms a SiLU epilogue can be built from on sm_100a: tanh.approx.f32, ex2.approx.f32,
// rcp.approx.f32, tanh.approx.bf16x2 / f16x2 (two results per instruction), and the whole SiLU pair used by the kernels
// (2 FFMA2 + 2 MUFU.TANH + F2FP); 8 independent chains per thread; results per clock per SM for 4 / 8 / 16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o build/mufu scripts/ubench/mufu.cu && build/mufu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[8];
    unsigned int H[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0.01f * (i + threadIdx.x); H[i] = 0x3c003c00u + i + threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
                else if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                else if (MODE == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                else if (MODE == 3) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(H[i]));
                else if (MODE == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(H[i]));
                else if (MODE == 5) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(H[i]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(H[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 1000;
    const char* names[6] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.bf16x2", "tanh.f16x2", "ex2.bf16x2"};
    for (int mode = 0; mode < 6; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k<0><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 1: k<1><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 2: k<2><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 3: k<3><<<148, warps * 32>>>(out, iters, cyc); break;
                    case 4: k<4><<<148, warps * 32>>>(out, iters, cyc); break;
                    default: k<5><<<148, warps * 32>>>(out, iters, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double instr = (double)iters * 32 * warps;                       // warp instructions per SM
            const double res = mode >= 3 ? 64 : 32;
            printf("%-16s %2d warps/SM: %8lld cycles, %.3f warp-instr/clk/SM, %.1f results/clk/SM\n", names[mode], warps, c,
                   instr / c, instr * res / c);
        }
    return 0;
}
