// Arithmetic of the head decode, shared by head_decode.cu (stand-alone kernels) and the decode-fused epilogue of the
// head's final 1x1 convolutions (conv_gemm.cu): both must produce the SAME bits from the same fp32 logits.
//
// Reference: DFL.forward (yolov8/model/components.py:176-191, 16 bins), the eval branch of Head.forward
// (yolov8/model/yolov8_head.py:127-144), Head.make_anchors (:146-158) and the candidate selection of the
// post-process (yolov8/tools/test.py:166-179).  All fp32.
#pragma once
#include "common.cuh"

namespace yms {

constexpr int kRegMax = 16;

// ex2/rcp approximations: |error| < 3e-6 on the score (stated tolerance of the decode tests: 5e-6)
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// xyxy exactly as tools/test.py:172-177 (w/2 is exact, so w*0.5f == w/2)
__device__ __forceinline__ float4 to_xyxy(float cx, float cy, float w, float h) {
    float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

// One DFL side: sum_k k * softmax(v)_k over the 16 bins.
__device__ __forceinline__ float dfl_expectation(const float (&v)[kRegMax]) {
    float mx = v[0];
    #pragma unroll
    for (int k = 1; k < kRegMax; ++k) mx = fmaxf(mx, v[k]);
    float sum = 0.f, wsum = 0.f;
    #pragma unroll
    for (int k = 0; k < kRegMax; ++k) { const float e = __expf(v[k] - mx); sum += e; wsum = fmaf((float)k, e, wsum); }
    return wsum / sum;
}

// (l, t, r, b) distances in grid units + anchor centre -> (cx, cy, w, h) in input pixels (yolov8_head.py:139-143).
__device__ __forceinline__ float4 dfl_box(float ax, float ay, float4 d, float st) {
    const float x1 = ax - d.x, y1 = ay - d.y, x2 = ax + d.z, y2 = ay + d.w;
    float4 box;
    box.x = ((x1 + x2) / 2.0f) * st;
    box.y = ((y1 + y2) / 2.0f) * st;
    box.z = (x2 - x1) * st;
    box.w = (y2 - y1) * st;
    return box;
}

// 16 class logits -> 16 sigmoid scores (r) + the chunk's first maximum (torch.max: lowest index wins on ties).
__device__ __forceinline__ void cls_chunk16(const float (&v)[16], int c0, float4 (&r)[4], float& best, int& bi) {
    best = -INFINITY; bi = 0x7fffffff;
    #pragma unroll
    for (int q = 0; q < 4; ++q) {
        r[q].x = sigmoid_f(v[4 * q]); r[q].y = sigmoid_f(v[4 * q + 1]); r[q].z = sigmoid_f(v[4 * q + 2]); r[q].w = sigmoid_f(v[4 * q + 3]);
        const float rr[4] = {r[q].x, r[q].y, r[q].z, r[q].w};
        #pragma unroll
        for (int j = 0; j < 4; ++j)
            if (rr[j] > best || bi == 0x7fffffff) { best = rr[j]; bi = c0 + 4 * q + j; }
    }
}

}  // namespace yms
