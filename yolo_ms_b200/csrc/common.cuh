// Shared helpers for libyms_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/yms_b200.h"

namespace yms {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- bf16 helpers --------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float silu_f(float x) {
    // x * sigmoid(x); ex2/rcp approximations are ~1e-6 relative, far below bf16 resolution
    return __fdividef(x, 1.0f + __expf(-x));
}

}  // namespace yms
