// CUDA-core kernels around the GEMM convolutions: SPPF pooling, nearest upsample, image resampling
// (the depthwise k x k kernel lives in ms_fused.cu, the stem in stem_tc.cu).  All activations are NHWC bf16 with an explicit pixel stride so that
// outputs land directly inside the channel slice of the concat buffer that consumes them.
#include "common.cuh"

#include <stdlib.h>

namespace yms {
namespace {

// =====================================================================================
// SPPF pooling (components.py:141-146).  Chained 5x5/s1/p2 max pools with -inf padding equal
// single 5x5, 9x9 and 13x13 windows clipped at the border; max is exact in bf16, so the
// results are bit-identical to the chained form.  One thread = one pixel x 8 channels (16 B);
// separable: vertical running max over rows is recomputed per thread from L1/L2-resident data
// (the map is 20x20..40x40 pixels).  Writes slots 1..3 of the concat buffer.
// =====================================================================================
__device__ __forceinline__ uint4 max_bf16x8(uint4 a, uint4 b) {
    uint4 r;
    __nv_bfloat162* ra = reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162* rb = reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162* rr = reinterpret_cast<__nv_bfloat162*>(&r);
    #pragma unroll
    for (int i = 0; i < 4; ++i) rr[i] = __hmax2(ra[i], rb[i]);
    return r;
}

__global__ void __launch_bounds__(256) sppf_pool_kernel(__nv_bfloat16* buf, long long ps, int h, int w, int c) {
    const int groups = c >> 3;
    const long long total = (long long)h * w * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const int px = (int)((idx / groups) % w);
    const int py = (int)(idx / ((long long)groups * w));
    __nv_bfloat16* base = buf + (size_t)b * h * w * ps + g * 8;
    const uint32_t ninf2 = 0xff80ff80u;            // bf16 -inf pair
    uint4 m5 = make_uint4(ninf2, ninf2, ninf2, ninf2), m9 = m5, m13 = m5;
    for (int dy = -6; dy <= 6; ++dy) {
        const int yy = py + dy;
        if (yy < 0 || yy >= h) continue;
        const int ady = dy < 0 ? -dy : dy;
        uint4 r5 = make_uint4(ninf2, ninf2, ninf2, ninf2), r9 = r5, r13 = r5;
        for (int dx = -6; dx <= 6; ++dx) {
            const int xx = px + dx;
            if (xx < 0 || xx >= w) continue;
            const int adx = dx < 0 ? -dx : dx;
            const uint4 v = *reinterpret_cast<const uint4*>(base + ((size_t)yy * w + xx) * ps);
            r13 = max_bf16x8(r13, v);
            if (adx <= 4) r9 = max_bf16x8(r9, v);
            if (adx <= 2) r5 = max_bf16x8(r5, v);
        }
        m13 = max_bf16x8(m13, r13);
        if (ady <= 4) m9 = max_bf16x8(m9, r9);
        if (ady <= 2) m5 = max_bf16x8(m5, r5);
    }
    __nv_bfloat16* o = base + ((size_t)py * w + px) * ps;
    *reinterpret_cast<uint4*>(o + c) = m5;
    *reinterpret_cast<uint4*>(o + 2 * c) = m9;
    *reinterpret_cast<uint4*>(o + 3 * c) = m13;
}

// Shared-memory version for maps up to 48x48: one CTA = one image x 8 * V channels (V = 2 when c % 16 == 0: the two 16-byte
// vectors of a pixel are read / written by adjacent threads, i.e. whole 32-byte sectors; with V = 1 every access moves half a
// sector).  The three chained 5x5 pools are computed as chained SEPARABLE passes (row max of 5, then column max of 5) on the
// plane held in shared memory: 30 smem reads per output instead of 169 global reads.
constexpr int kPoolMaxHW = 48 * 48;
template <int V>
__global__ void __launch_bounds__(256) sppf_pool_smem_kernel(__nv_bfloat16* buf, long long ps, int h, int w, int c) {
    extern __shared__ uint4 s_pool[];
    const int g = blockIdx.x, b = blockIdx.y;
    const int hw = h * w, n = hw * V;                              // item i = (pixel i / V, vector i % V)
    uint4* s_a = s_pool;
    uint4* s_t = s_pool + n;
    __nv_bfloat16* base = buf + (size_t)b * hw * ps + g * (8 * V);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_a[i] = *reinterpret_cast<const uint4*>(base + (size_t)(i / V) * ps + (i % V) * 8);
    __syncthreads();
    for (int level = 1; level <= 3; ++level) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {        // row pass: s_t = max over x-2..x+2 of s_a
            const int px = i / V, y = px / w, x = px - y * w;
            uint4 m = s_a[i];
            #pragma unroll
            for (int d = 1; d <= 2; ++d) {
                if (x - d >= 0) m = max_bf16x8(m, s_a[i - d * V]);
                if (x + d < w) m = max_bf16x8(m, s_a[i + d * V]);
            }
            s_t[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {        // column pass: s_a = max over y-2..y+2 of s_t
            const int px = i / V, y = px / w;
            uint4 m = s_t[i];
            #pragma unroll
            for (int d = 1; d <= 2; ++d) {
                if (y - d >= 0) m = max_bf16x8(m, s_t[i - d * w * V]);
                if (y + d < h) m = max_bf16x8(m, s_t[i + d * w * V]);
            }
            s_a[i] = m;
            *reinterpret_cast<uint4*>(base + (size_t)px * ps + (size_t)level * c + (i % V) * 8) = m;
        }
        __syncthreads();
    }
}

// =====================================================================================
// Nearest x2 upsample (components.py:159-160) into a channel slice.  One thread = one INPUT
// pixel x 8 channels: one 16 B load, four 16 B stores.
// =====================================================================================
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, long long xps,
                                                         int h, int w, int c, __nv_bfloat16* __restrict__ y, long long yps) {
    const int groups = c >> 3;
    const long long total = (long long)h * w * groups;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (idx >= total) return;
    const int g = (int)(idx % groups);
    const int px = (int)((idx / groups) % w);
    const int py = (int)(idx / ((long long)groups * w));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * h + py) * w + px) * xps + g * 8));
    __nv_bfloat16* o = y + ((size_t)(b * 2 * h + 2 * py) * (2 * w) + 2 * px) * yps + g * 8;
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + yps) = v;
    *reinterpret_cast<uint4*>(o + (size_t)2 * w * yps) = v;
    *reinterpret_cast<uint4*>(o + (size_t)2 * w * yps + yps) = v;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace
}  // namespace yms

using namespace yms;

int yms_stem_tc_launch(const float* x, const unsigned char* xu8, const float* mean, const float* stdv, int batch, int in_h, int in_w,
                       int c_out, const float* weight, const float* bias, void* y, int64_t y_ps, cudaStream_t stream);   // stem_tc.cu

extern "C" int yms_stem_conv(const float* x, int batch, int in_h, int in_w, int c_out, const float* weight,
                             const float* bias, void* y, int64_t y_ps, void* stream) {
    if (batch <= 0 || in_h <= 0 || in_w <= 0 || (in_h & 1) || (in_w & 1)) return fail(YMS_E_ARG, "stem: bad image size");
    if (c_out <= 0 || (c_out % 16) != 0 || c_out > 128) return fail(YMS_E_UNSUPPORTED, "stem: c_out must be a multiple of 16 (<= 128)");
    if (!x || !weight || !bias || !y || !aligned16(y) || (y_ps % 8) != 0) return fail(YMS_E_ARG, "stem: bad pointers/strides");
    return yms_stem_tc_launch(x, nullptr, nullptr, nullptr, batch, in_h, in_w, c_out, weight, bias, y, y_ps, (cudaStream_t)stream);
}

extern "C" int yms_stem_conv_u8(const uint8_t* x, int batch, int in_h, int in_w, int c_out, const float* weight,
                                const float* bias, const float* host_mean, const float* host_std, void* y, int64_t y_ps, void* stream) {
    if (batch <= 0 || in_h <= 0 || in_w <= 0 || (in_h & 1) || (in_w & 1)) return fail(YMS_E_ARG, "stem_u8: bad image size");
    if (c_out <= 0 || (c_out % 16) != 0 || c_out > 128) return fail(YMS_E_UNSUPPORTED, "stem_u8: c_out must be a multiple of 16 (<= 128)");
    if (!x || !weight || !bias || !y || !host_mean || !host_std || !aligned16(y) || (y_ps % 8) != 0) return fail(YMS_E_ARG, "stem_u8: bad pointers/strides");
    for (int c = 0; c < 3; ++c) if (!(host_std[c] > 0.f)) return fail(YMS_E_ARG, "stem_u8: std must be positive");
    return yms_stem_tc_launch(nullptr, x, host_mean, host_std, batch, in_h, in_w, c_out, weight, bias, y, y_ps, (cudaStream_t)stream);
}

extern "C" int yms_sppf_pool(void* buf, int64_t ps, int batch, int h, int w, int c, void* stream) {
    if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8) != 0) return fail(YMS_E_ARG, "sppf: bad sizes (c % 8 == 0)");
    if (!buf || !aligned16(buf) || (ps % 8) != 0 || ps < 4 * (int64_t)c) return fail(YMS_E_ARG, "sppf: bad buffer");
    if (h * w <= kPoolMaxHW) {
        const int v = (c % 16 == 0 && 2 * h * w <= kPoolMaxHW) ? 2 : 1;
        dim3 grid(c / (8 * v), batch);
        const size_t smem = (size_t)2 * v * h * w * sizeof(uint4);
        static std::atomic<unsigned long long> attr_seen{0};
        if (first_use_on_device(attr_seen)) {
            cudaError_t e = cudaFuncSetAttribute(sppf_pool_smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPoolMaxHW * (int)sizeof(uint4));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(sppf_pool_smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPoolMaxHW * (int)sizeof(uint4));
            if (e != cudaSuccess) return fail((int)e, "sppf: smem attribute");
        }
        if (v == 2) sppf_pool_smem_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(buf), ps, h, w, c);
        else sppf_pool_smem_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(buf), ps, h, w, c);
        return check_launch("sppf_pool_smem_kernel");
    }
    long long total = (long long)h * w * (c / 8);
    dim3 grid((unsigned)((total + 255) / 256), batch);
    sppf_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(buf), ps, h, w, c);
    return check_launch("sppf_pool_kernel");
}

extern "C" int yms_upsample2x(const void* x, int64_t xps, int batch, int h, int w, int c, void* y, int64_t yps, void* stream) {
    if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8) != 0) return fail(YMS_E_ARG, "upsample: bad sizes (c % 8 == 0)");
    if (!x || !y || !aligned16(x) || !aligned16(y) || (xps % 8) != 0 || (yps % 8) != 0) return fail(YMS_E_ARG, "upsample: bad pointers/strides");
    long long total = (long long)h * w * (c / 8);
    dim3 grid((unsigned)((total + 255) / 256), batch);
    upsample2x_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), xps, h, w, c,
                                                             reinterpret_cast<__nv_bfloat16*>(y), yps);
    return check_launch("upsample2x_kernel");
}

// ---------------------------------------------------------------------------------------
// Pillow-exact fixed-point resampling pass (T.Resize of the reference's pre-processing,
// yolov8/tools/test.py:114-119, runs Image.resize(BILINEAR) = Pillow's src/libImaging/Resample.c).
// One thread per output byte: acc = 2^21 + sum_k src * coeff (22-bit fixed point), out = clip8(acc >> 22).
// ---------------------------------------------------------------------------------------
namespace yms {
namespace {
__global__ void __launch_bounds__(256) resample_u8_kernel(const unsigned char* __restrict__ src, long long src_rs, unsigned char* __restrict__ dst,
                                                          long long dst_rs, int dst_h, int dst_wc, int channels,
                                                          const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize, int horizontal) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;         // byte column of the output row (x * channels + c)
    const int row = blockIdx.y;
    if (col >= dst_wc || row >= dst_h) return;
    int acc = 1 << 21;
    if (horizontal) {
        const int xo = col / channels, c = col - xo * channels;
        const int lo = bounds[2 * xo], cnt = bounds[2 * xo + 1];
        const int* k = coeffs + (size_t)xo * ksize;
        const unsigned char* s = src + (size_t)row * src_rs + (size_t)lo * channels + c;
        for (int i = 0; i < cnt; ++i) acc += (int)s[(size_t)i * channels] * __ldg(k + i);
    } else {
        const int lo = bounds[2 * row], cnt = bounds[2 * row + 1];
        const int* k = coeffs + (size_t)row * ksize;
        const unsigned char* s = src + (size_t)lo * src_rs + col;
        for (int i = 0; i < cnt; ++i) acc += (int)s[(size_t)i * src_rs] * __ldg(k + i);
    }
    acc >>= 22;
    dst[(size_t)row * dst_rs + col] = (unsigned char)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
}
}  // namespace
}  // namespace yms

extern "C" int yms_resample_u8(const uint8_t* src, int src_h, int src_w, int channels, int64_t src_row_stride,
                               uint8_t* dst, int dst_h, int dst_w, int64_t dst_row_stride,
                               const int32_t* bounds, const int32_t* coeffs, int ksize, int horizontal, void* stream) {
    if (!src || !dst || !bounds || !coeffs) return fail(YMS_E_ARG, "resample: null pointer");
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || channels <= 0 || ksize <= 0) return fail(YMS_E_ARG, "resample: bad sizes");
    if (horizontal ? (dst_h != src_h) : (dst_w != src_w)) return fail(YMS_E_ARG, "resample: a pass changes one axis only");
    if (src_row_stride < (int64_t)src_w * channels || dst_row_stride < (int64_t)dst_w * channels) return fail(YMS_E_ARG, "resample: bad row stride");
    if (dst_h > 65535) return fail(YMS_E_UNSUPPORTED, "resample: more than 65535 output rows");
    dim3 grid(ceil_div(dst_w * channels, 256), dst_h);
    resample_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_row_stride, dst, dst_row_stride, dst_h, dst_w * channels, channels,
                                                                bounds, coeffs, ksize, horizontal);
    return check_launch("resample_u8_kernel");
}
