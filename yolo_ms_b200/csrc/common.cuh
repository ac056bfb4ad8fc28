// Shared helpers for libyms_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/yms_b200.h"

namespace yms {

// Debug / experiment switches of the library, set through yms_debug_set_option (capi.cu) -- never read from the environment,
// never consulted per launch beyond a plain load.
struct DebugOptions {
    int pdl_off;            // 1: launch without programmatic stream serialization
    int conv_half;          // 1: "half" CTAs for N <= 128 plans of the generic conv kernel (two CTAs per SM)
    int stem_gather;        // 1: force the gather stem kernel (bit-identical to the TMA-fed one)
    int nms_groups;         // > 0: CTAs per image of the NMS kernel
    int nms_mask_tiles;     // IoU-bitmask path limit (default 8)
    int nms_poll_ns;        // back-off of the pipelined greedy path (default 256)
    int nms_sort_bitonic;   // 1: single bitonic sort instead of per-segment sorts
};
extern DebugOptions g_opt;

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

// cudaFuncSetAttribute is per DEVICE: returns true the first time it is called on the current device for this flag word
// (one word per call site), so a second GPU driven from the same process gets its dynamic shared-memory opt-in too.
inline bool first_use_on_device(std::atomic<unsigned long long>& seen) {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    return !(seen.fetch_or(bit) & bit);
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
