// Micro-benchmark (B200): what does one k-block cost in the tcgen05 issue loop of the conv kernels?
// One issuer warp per CTA, operands = garbage in shared memory (valid SW128 K-major descriptors), no producer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -I yolo_ms_b200/csrc scripts/ubench/mma_issue.cu -o build/mma_issue
// Output: cycles per k-block for (N, MMAs per k-block, #independent accumulators, variant).
#include "tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace yms::tc;

struct P { int iters; long long* out; };
// variant bits: 1 = try_wait on an already-complete barrier per k-block, 2 = tcgen05.fence::after, 4 = elect inside the loop
//               (else the whole loop runs in one elected lane), 8 = __syncwarp per k-block, 16 = prefetch the try_wait of the next k-block
template <int kN, int kMmas, int kAccs, int kVariant>
__global__ void __launch_bounds__(128, 1) k(const P p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[4];
    __shared__ uint32_t tslot;
    const uint32_t base = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t adesc = make_sw128_desc(base), bdesc = make_sw128_desc(base + 16384);
        const uint32_t done = smem_u32(&bars[0]), ready = smem_u32(&bars[1]), fin = smem_u32(&bars[2]);
        const bool inner = kVariant & 4;
        long long t0 = 0;
        if (inner || elect_one()) {
            t0 = clock64();
            uint32_t ok_next = 1;
            for (int it = 0; it < p.iters; ++it) {
                if (kVariant & 16) { if (!ok_next) mbar_wait(ready, 1u); }
                else if (kVariant & 1) mbar_wait(ready, 1u);           // fresh barrier: parity-1 wait succeeds immediately
                if (kVariant & 2) tc_fence_after();
                if (kVariant & 16) ok_next = mbar_try_wait(ready, 1u);
                if (!inner || elect_one()) {
                    #pragma unroll
                    for (int m = 0; m < kMmas; ++m) {
                        constexpr int dummy = 0; (void)dummy;
                        umma_bf16(tmem + (m % kAccs) * kN, adesc + 2 * (m & 3), bdesc + 2 * (m & 3), idesc, (m >= kAccs) ? 1u : (it ? 1u : 0u));
                    }
                    umma_commit(done);
                }
                if (kVariant & 8) __syncwarp();
            }
            if (!inner || elect_one()) umma_commit(fin);
            if (inner) __syncwarp();
            mbar_wait(fin, 0u);
            const long long t1 = clock64();
            if (blockIdx.x == 0 && (inner ? lane == 0 : true)) p.out[0] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}


template <int kN, int kMmas, int kAccs, int kVariant>
void run(long long* out) {
    if (kAccs * kN > 512) return;
    const int iters = 2000;
    cudaFuncSetAttribute(k<kN, kMmas, kAccs, kVariant>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    P p{iters, out};
    k<kN, kMmas, kAccs, kVariant><<<148, 128, 64 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("%d %d %d %d %.1f %.1f\n", kN, kMmas, kAccs, kVariant, (double)c / iters, (double)c / iters / kMmas);
}
template <int kN, int kVariant> void run_n(long long* out) {
    run<kN, 2, 1, kVariant>(out); run<kN, 2, 2, kVariant>(out); run<kN, 4, 1, kVariant>(out); run<kN, 4, 2, kVariant>(out);
    run<kN, 8, 1, kVariant>(out); run<kN, 8, 2, kVariant>(out); run<kN, 36, 1, kVariant>(out); run<kN, 36, 2, kVariant>(out); run<kN, 36, 4, kVariant>(out);
}
template <int kVariant> void run_v(long long* out) { run_n<32, kVariant>(out); run_n<64, kVariant>(out); run_n<128, kVariant>(out); run_n<256, kVariant>(out); }
int main() {
    long long* out; cudaMalloc(&out, 8);
    printf("N mmas naccs variant cyc_per_kblock cyc_per_mma\n");
    run_v<0>(out); run_v<3>(out); run_v<15>(out); run_v<30>(out);
    return 0;
}
