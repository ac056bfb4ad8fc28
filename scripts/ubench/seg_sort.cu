// Micro-benchmark (B200), prepared for the next round -- NOT yet run on hardware, NOT part of the product library.
// Question: how much of nms_kernel's sort phase (54 kcycles for ~2100 keys in ~20 class segments per CTA, shared-memory-
// bandwidth-bound even with one warp per segment) goes away if segments of <= 256 keys are sorted in
// REGISTERS -- 8 keys per lane, compare-exchange partners reached with __shfl_xor_sync -- instead of in shared memory?
//   variant 0: nms.cu::segment_sort<false> (shared memory, one warp per segment, normalised bitonic network)
//   variant 1: warp_sort_regs<E> (E = 1, 2, 4, 8 keys per lane; same network, strides >= E cross lanes by shuffle)
// Both sort the same segments; the result is checked against std::sort; cycles are the CTA's clock64() span.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared scripts/ubench/seg_sort.cu -o build/seg_sort && build/seg_sort
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned long long u64;
constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;

// ---- variant 0: the shared-memory network of nms.cu -------------------------------------------------------------
__device__ __forceinline__ void segment_sort_smem(u64* s, int len, int t0, int nthreads) {
    if (len < 2) return;
    int P = 2;
    while (P < len) P <<= 1;
    const int half_pairs = P >> 1;
    for (int k = 2, lg = 0; k <= P; k <<= 1, ++lg) {
        for (int t = t0; t < half_pairs; t += nthreads) {
            const int blk = t >> lg, off = t & ((k >> 1) - 1);
            const int i = blk * k + off, l = blk * k + (k - 1 - off);
            if (l < len) { const u64 x = s[i], y = s[l]; if (x > y) { s[i] = y; s[l] = x; } }
        }
        __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = t0; t < half_pairs; t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                if (l < len) { const u64 x = s[i], y = s[l]; if (x > y) { s[i] = y; s[l] = x; } }
            }
            __syncwarp();
        }
    }
}

// ---- variant 1: registers + shuffles ---------------------------------------------------------------------------------
// Element index = lane * E + r.  XOR partners below E stay in the lane; a half-cleaner of stride j >= E pairs lane with
// lane ^ (j / E) (same r); the mirror step of merge size k > E pairs (lane, r) with (lane ^ (k / E - 1), E - 1 - r).
// The lower index keeps the minimum.  Positions >= len hold +inf (all ones) and therefore never move.
__device__ __forceinline__ void cmpx(u64& a, u64& b) { if (a > b) { const u64 t = a; a = b; b = t; } }

template <int E>
__device__ __forceinline__ void warp_sort_regs(u64* s, int len, int lane) {
    u64 v[E];
    #pragma unroll
    for (int r = 0; r < E; ++r) { const int i = lane * E + r; v[r] = (i < len) ? s[i] : ~0ull; }
    constexpr int P = 32 * E;
    #pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
        if (k <= E) {                                              // mirror inside the lane
            #pragma unroll
            for (int r = 0; r < E; ++r) { const int q = r ^ (k - 1); if (q > r) cmpx(v[r], v[q]); }
        } else {                                                   // mirror across lanes
            const int mm = k / E - 1;
            const bool lower = (lane & ((mm + 1) >> 1)) == 0;
            u64 o[E];
            #pragma unroll
            for (int r = 0; r < E; ++r) o[r] = __shfl_xor_sync(0xffffffffu, v[E - 1 - r], mm);
            #pragma unroll
            for (int r = 0; r < E; ++r) v[r] = lower ? (v[r] < o[r] ? v[r] : o[r]) : (v[r] > o[r] ? v[r] : o[r]);
        }
        #pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j < E) {
                #pragma unroll
                for (int r = 0; r < E; ++r) if ((r & j) == 0) cmpx(v[r], v[r | j]);
            } else {
                const int m = j / E;
                const bool lower = (lane & m) == 0;
                #pragma unroll
                for (int r = 0; r < E; ++r) {
                    const u64 o = __shfl_xor_sync(0xffffffffu, v[r], m);
                    v[r] = lower ? (v[r] < o ? v[r] : o) : (v[r] > o ? v[r] : o);
                }
            }
        }
    }
    #pragma unroll
    for (int r = 0; r < E; ++r) { const int i = lane * E + r; if (i < len) s[i] = v[r]; }
}

__device__ __forceinline__ void warp_sort_dispatch(u64* s, int len, int lane) {
    if (len < 2) return;
    if (len <= 32) warp_sort_regs<1>(s, len, lane);
    else if (len <= 64) warp_sort_regs<2>(s, len, lane);
    else if (len <= 128) warp_sort_regs<4>(s, len, lane);
    else warp_sort_regs<8>(s, len, lane);                          // <= 256
}

template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1) sort_kernel(const u64* in, u64* out, const int* seg_start, int nseg, int total, long long* cycles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* s = reinterpret_cast<u64*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < total; i += kThreads) s[i] = in[(size_t)blockIdx.x * total + i];
    __syncthreads();
    const long long t0 = clock64();
    for (int c = warp; c < nseg; c += kWarps) {
        const int s0 = seg_start[c], len = seg_start[c + 1] - s0;
        if (kVariant == 0) segment_sort_smem(s + s0, len, lane, 32);
        else warp_sort_dispatch(s + s0, len, lane);
    }
    __syncthreads();
    const long long t1 = clock64();
    for (int i = tid; i < total; i += kThreads) out[(size_t)blockIdx.x * total + i] = s[i];
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    // the bench workload's shape: ~2100 keys per CTA in 20 class segments of very different sizes (all <= 256 here: larger
    // classes keep the CTA-wide path), 128 CTAs
    const int nseg = 24, ctas = 128;
    std::vector<int> start(nseg + 1, 0);
    srand(7);
    for (int c = 0; c < nseg; ++c) start[c + 1] = start[c] + 1 + rand() % ((c % 5 == 0) ? 256 : (c % 3 == 0 ? 130 : 60));
    const int total = start[nseg];
    std::vector<u64> h((size_t)ctas * total);
    for (auto& x : h) x = ((u64)rand() << 40) ^ ((u64)rand() << 20) ^ (u64)rand();
    u64 *din, *dout; int* dstart; long long* dcyc;
    cudaMalloc(&din, h.size() * 8); cudaMalloc(&dout, h.size() * 8); cudaMalloc(&dstart, (nseg + 1) * 4); cudaMalloc(&dcyc, ctas * 8);
    cudaMemcpy(din, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dstart, start.data(), (nseg + 1) * 4, cudaMemcpyHostToDevice);
    std::vector<u64> want = h;
    for (int b = 0; b < ctas; ++b)
        for (int c = 0; c < nseg; ++c) std::sort(want.begin() + (size_t)b * total + start[c], want.begin() + (size_t)b * total + start[c + 1]);
    const size_t smem = (size_t)total * 8;
    cudaFuncSetAttribute(sort_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(sort_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("%d keys per CTA in %d segments, %d CTAs\n", total, nseg, ctas);
    for (int variant = 0; variant < 2; ++variant) {
        for (int rep = 0; rep < 3; ++rep) {
            if (variant == 0) sort_kernel<0><<<ctas, kThreads, smem>>>(din, dout, dstart, nseg, total, dcyc);
            else sort_kernel<1><<<ctas, kThreads, smem>>>(din, dout, dstart, nseg, total, dcyc);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<u64> got(h.size());
        std::vector<long long> cyc(ctas);
        cudaMemcpy(got.data(), dout, got.size() * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(cyc.data(), dcyc, ctas * 8, cudaMemcpyDeviceToHost);
        long long mx = 0, sum = 0;
        for (long long c : cyc) { mx = std::max(mx, c); sum += c; }
        printf("variant %d (%s): %s, cycles per CTA mean %.0f max %lld\n", variant, variant ? "registers + shuffles" : "shared memory",
               got == want ? "sorted OK" : "WRONG RESULT", (double)sum / ctas, mx);
    }
    return 0;
}
