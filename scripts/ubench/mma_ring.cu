// Micro-benchmark (B200): producer/consumer ring between a "TMA" warp (arrives only, no data movement) and the
// tcgen05 issuer, in the two issue-loop styles considered for the conv kernels.
//   style 0: whole warp loops, elect_one() inside per k-block, __syncwarp (current conv_gemm.cu)
//   style 1: ONE elected thread runs the whole loop
//   style 2: style 1 + the try_wait of the next stage is issued before the MMAs of the current one
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -I yolo_ms_b200/csrc scripts/ubench/mma_ring.cu -o build/mma_ring
#include "tc_ptx.cuh"
#include <cstdio>
#include <cstdlib>
using namespace yms::tc;

struct P { int iters, block_n, stages, ksteps, kb_per_tile; long long* out; };

template <int kStyle>
__global__ void __launch_bounds__(128, 1) k(const P p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[32];
    __shared__ uint32_t tslot;
    const uint32_t base = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t fin = bar0 + 8u * 16, tfull = bar0 + 8u * 17;
    if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(bar0 + 8u * i, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(smem_u32(&tslot), 512);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tslot;
    const int stage_bytes = 16384 + p.block_n * 128;
    if (warp == 0) {
        if (elect_one()) {                         // "producer": waits for the slot, then just arrives
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < p.iters; ++it) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive(full_bar(stage));
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        long long t0 = clock64();
        if (kStyle == 0) {
            int stage = 0; uint32_t phase = 0; int kbi = 0;
            for (int it = 0; it < p.iters; ++it) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const int ksteps = p.ksteps;
                const uint32_t sa = base + stage * stage_bytes;
                const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + 16384);
                const uint32_t first = kbi ? 1u : 0u;
                const bool last = (kbi == p.kb_per_tile - 1);
                if (elect_one()) {
                    #pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        if (kk < ksteps) umma_bf16(tmem, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, kk ? 1u : first);
                    umma_commit(empty_bar(stage));
                    if (last) umma_commit(tfull);
                }
                __syncwarp();
                if (++kbi == p.kb_per_tile) kbi = 0;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (elect_one()) umma_commit(fin);
            __syncwarp();
            mbar_wait(fin, 0u);
        } else if (elect_one()) {
            int stage = 0; uint32_t phase = 0; int kbi = 0;
            const uint64_t hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            uint32_t a16 = (base & 0x3FFFFu) >> 4;
            const uint32_t st16 = (uint32_t)stage_bytes >> 4;
            uint32_t ok = (kStyle == 2) ? mbar_try_wait(full_bar(0), 0u) : 0u;
            for (int it = 0; it < p.iters; ++it) {
                if (kStyle == 2) { if (!ok) mbar_wait(full_bar(stage), phase); }
                else mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                int nstage = stage + 1; uint32_t nphase = phase;
                if (nstage == p.stages) { nstage = 0; nphase ^= 1u; }
                if (kStyle == 2) ok = mbar_try_wait(full_bar(nstage), nphase);
                const uint32_t first = kbi ? 1u : 0u;
                #pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    if (kk < p.ksteps) umma_bf16(tmem, hi | (uint64_t)(a16 + 2 * kk), hi | (uint64_t)(a16 + 1024 + 2 * kk), idesc, kk ? 1u : first);
                umma_commit(empty_bar(stage));
                if (kbi == p.kb_per_tile - 1) { umma_commit(tfull); kbi = 0; } else ++kbi;
                a16 += st16; if (nstage == 0) a16 = (base & 0x3FFFFu) >> 4;
                stage = nstage; phase = nphase;
            }
            umma_commit(fin);
            mbar_wait(fin, 0u);
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) p.out[0] = t1 - t0;
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int kStyle> void run(long long* out, int n, int ksteps, int kbt, int stages = 4) {
    cudaFuncSetAttribute(k<kStyle>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    P p{4000, n, stages, ksteps, kbt, out};
    k<kStyle><<<148, 128, 200 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("style %d N %d ksteps %d kb/tile %d stages %d: %.1f cycles per k-block\n", kStyle, n, ksteps, kbt, stages, (double)c / p.iters);
}
int main() {
    long long* out; cudaMalloc(&out, 8);
    for (int st : {1, 2, 4, 8})
        for (int n : {64, 256})
            for (int ks : {2, 4}) { run<0>(out, n, ks, 9, st); run<1>(out, n, ks, 9, st); run<2>(out, n, ks, 9, st); }
    return 0;
}
