// CUDA-core kernels around the GEMM convolutions: stem conv, SPPF pooling, nearest upsample,
// depthwise k x k.  All activations are NHWC bf16 with an explicit pixel stride so that
// outputs land directly inside the channel slice of the concat buffer that consumes them.
#include "common.cuh"

#include <stdlib.h>

namespace yms {
namespace {

// =====================================================================================
// Stem: backbone.conv0 (yolov8_backbone.py:39), 3x3 stride 2 pad 1 on the NCHW fp32 image,
// folded BN + SiLU, NHWC bf16 out.  One thread = one output pixel x CO_T output channels.
// The 27 input taps are read once into registers; weights are broadcast from shared memory.
// =====================================================================================
template <int CO_T>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, int in_h, int in_w, int c_out,
                                                        const float* __restrict__ wgt, const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ y, long long y_ps) {
    extern __shared__ float s_w[];                 // [27][c_out] then bias[c_out]
    const int out_h = in_h >> 1, out_w = in_w >> 1;
    for (int i = threadIdx.x; i < 27 * c_out; i += blockDim.x) {
        // wgt is [c_out][ci][ky][kx]; store as [(ci*9+ky*3+kx)][c_out]
        int co = i % c_out, t = i / c_out;
        s_w[i] = wgt[co * 27 + t];
    }
    float* s_b = s_w + 27 * c_out;
    for (int i = threadIdx.x; i < c_out; i += blockDim.x) s_b[i] = bias[i];
    __syncthreads();

    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    const int b = blockIdx.z;
    if (ox >= out_w) return;
    float in[27];
    const float* xb = x + (size_t)b * 3 * in_h * in_w;
    #pragma unroll
    for (int ci = 0; ci < 3; ++ci)
        #pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy + ky - 1;
            #pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox + kx - 1;
                float v = 0.f;
                if (iy >= 0 && iy < in_h && ix >= 0 && ix < in_w) v = __ldg(xb + ((size_t)ci * in_h + iy) * in_w + ix);
                in[ci * 9 + ky * 3 + kx] = v;
            }
        }
    __nv_bfloat16* yo = y + ((size_t)(b * out_h + oy) * out_w + ox) * y_ps;
    for (int c0 = 0; c0 < c_out; c0 += CO_T) {
        float acc[CO_T];
        #pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[j] = s_b[c0 + j];
        #pragma unroll
        for (int t = 0; t < 27; ++t) {
            const float v = in[t];
            const float* wr = s_w + t * c_out + c0;
            #pragma unroll
            for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(v, wr[j], acc[j]);
        }
        uint32_t pk[CO_T / 2];
        #pragma unroll
        for (int j = 0; j < CO_T / 2; ++j) pk[j] = pack_bf16x2(silu_f(acc[2 * j]), silu_f(acc[2 * j + 1]));
        #pragma unroll
        for (int j = 0; j < CO_T / 8; ++j)
            *reinterpret_cast<uint4*>(yo + c0 + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
}

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

// Shared-memory version for maps up to 48x48: one CTA = one image x 8 channels.  The three chained
// 5x5 pools are computed as chained SEPARABLE passes (row max of 5, then column max of 5) on the
// plane held in shared memory: 30 smem reads per output instead of 169 global reads.
constexpr int kPoolMaxHW = 48 * 48;
__global__ void __launch_bounds__(256) sppf_pool_smem_kernel(__nv_bfloat16* buf, long long ps, int h, int w, int c) {
    extern __shared__ uint4 s_pool[];
    const int g = blockIdx.x, b = blockIdx.y;
    const int hw = h * w;
    uint4* s_a = s_pool;
    uint4* s_t = s_pool + hw;
    __nv_bfloat16* base = buf + (size_t)b * hw * ps + g * 8;
    for (int i = threadIdx.x; i < hw; i += blockDim.x) s_a[i] = *reinterpret_cast<const uint4*>(base + (size_t)i * ps);
    __syncthreads();
    for (int level = 1; level <= 3; ++level) {
        for (int i = threadIdx.x; i < hw; i += blockDim.x) {       // row pass: s_t = max over x-2..x+2 of s_a
            const int y = i / w, x = i - y * w;
            uint4 m = s_a[i];
            #pragma unroll
            for (int d = 1; d <= 2; ++d) {
                if (x - d >= 0) m = max_bf16x8(m, s_a[i - d]);
                if (x + d < w) m = max_bf16x8(m, s_a[i + d]);
            }
            s_t[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < hw; i += blockDim.x) {       // column pass: s_a = max over y-2..y+2 of s_t
            const int y = i / w;
            uint4 m = s_t[i];
            #pragma unroll
            for (int d = 1; d <= 2; ++d) {
                if (y - d >= 0) m = max_bf16x8(m, s_t[i - d * w]);
                if (y + d < h) m = max_bf16x8(m, s_t[i + d * w]);
            }
            s_a[i] = m;
            *reinterpret_cast<uint4*>(base + (size_t)i * ps + (size_t)level * c) = m;
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

// =====================================================================================
// Depthwise k x k (stride 1, pad k/2) + folded BN + SiLU, NHWC bf16 (MS-Block branches).
// A CTA computes a TH x TW output tile for 8*CG channels: the (TH+k-1) x (TW+k-1) halo tile
// is staged in shared memory with 16-byte loads, weights in shared memory as fp32, fp32
// accumulate.  One thread = one output pixel x 8 channels.
// =====================================================================================
// Depthwise k x k (+ folded BN bias + SiLU), NHWC bf16.  Persistent CTAs (two per SM) walk 32 x 8 pixel x 64 channel
// tiles; the halo tile of the NEXT work item is staged in shared memory with cp.async (16-byte copies in which 8 consecutive
// threads fetch the 128 contiguous bytes of one pixel, zero-filled outside the image = the conv padding) while the current
// one is computed, so loads and FMAs overlap inside a CTA.  A thread owns one 8-channel group and 4 consecutive output
// pixels of a row (two rows per thread): each staged input vector is converted to fp32 once and feeds up to 4 outputs
// (sliding window), the k weights of the current kernel row sit in registers.  k = 3 is memory-bound; k >= 5 is bound by
// the fp32 FMA pipe (2*k*k flops per element).  (The first version gave every CTA 16 channels, i.e. 32-byte fragments of
// each 128-byte line, and a load -> store chain per element: 1.5 TB/s.)
constexpr int kDwTW = 32, kDwOX = 4, kDwThreads = 256;
template <int K> struct DwTile { static constexpr int kTH = (K == 3) ? 8 : 4; };   // k >= 5: smaller tiles, two CTAs per SM still fit

struct DwArgs {
    const __nv_bfloat16* x; long long xps; int batch, h, w, c;
    const float* wgt; const float* bias; __nv_bfloat16* y; long long yps;
    int tiles_x, tiles_y, cblocks, total;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;                       // src-size 0: 16 bytes of zeros, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ float silu_tanh(float x) {   // same single-MUFU form as the conv epilogue
    const float hx = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(hx));
    return fmaf(hx, t, hx);
}

template <int K>
__global__ void __launch_bounds__(kDwThreads, 2) dwconv_kernel(const DwArgs a) {
    constexpr int kDwTH = DwTile<K>::kTH;
    constexpr int R = K / 2, IW = kDwTW + K - 1, IH = kDwTH + K - 1, kTile = IH * IW * 8;
    extern __shared__ __align__(16) unsigned char dw_smem[];
    uint4* s_in = reinterpret_cast<uint4*>(dw_smem);                              // 2 x [IH][IW][8]
    float* s_w = reinterpret_cast<float*>(dw_smem + (size_t)2 * kTile * 16);      // [K*K][64]
    float* s_b = s_w + K * K * 64;                                                // [64]
    const int tid = threadIdx.x;
    const int fg = tid & 7;
    const int per_c = a.tiles_x * a.tiles_y * a.batch;

    auto decode = [&](int t, int& cb, int& b, int& ty0, int& tx0) {
        cb = t / per_c; int r = t - cb * per_c;
        b = r / (a.tiles_x * a.tiles_y); r -= b * (a.tiles_x * a.tiles_y);
        ty0 = (r / a.tiles_x) * kDwTH; tx0 = (r % a.tiles_x) * kDwTW;
    };
    auto prefetch = [&](int t, int buf) {
        int cb, b, ty0, tx0; decode(t, cb, b, ty0, tx0);
        const int c0 = cb * 64;
        const bool gv = c0 + fg * 8 < a.c;
        const __nv_bfloat16* xb = a.x + (size_t)b * a.h * a.w * a.xps + c0 + fg * 8;
        const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_in + (size_t)buf * kTile);
        for (int i = tid; i < kTile; i += kDwThreads) {
            const int pix = i >> 3;
            const int r = pix / IW, q = pix - r * IW;
            const int yy = ty0 + r - R, xx = tx0 + q - R;
            const bool ok = gv && yy >= 0 && yy < a.h && xx >= 0 && xx < a.w;
            cp_async16(dst0 + (uint32_t)i * 16u, ok ? (const void*)(xb + ((size_t)yy * a.w + xx) * a.xps) : (const void*)a.x, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int t = blockIdx.x, it = 0, cur_cb = -1;
    if (t < a.total) prefetch(t, 0);
    for (; t < a.total; t += gridDim.x, ++it) {
        const int tn = t + gridDim.x;
        if (tn < a.total) { prefetch(tn, (it + 1) & 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        int cb, b, ty0, tx0; decode(t, cb, b, ty0, tx0);
        const int c0 = cb * 64;
        if (cb != cur_cb) {                                                       // tiles are ordered by channel block: rare
            __syncthreads();
            for (int i = tid; i < K * K * 64; i += kDwThreads) {
                const int ch = i & 63, tap = i >> 6;
                s_w[i] = (c0 + ch < a.c) ? a.wgt[(size_t)tap * a.c + c0 + ch] : 0.f;
            }
            if (tid < 64) s_b[tid] = (c0 + tid < a.c) ? a.bias[c0 + tid] : 0.f;
            cur_cb = cb;
        }
        __syncthreads();
        const uint4* tile = s_in + (size_t)(it & 1) * kTile;
        const int g = fg, xblk = (tid >> 3) & 7, yr = tid >> 6;                   // 8 groups x 8 x-blocks x 4 rows
        if (c0 + g * 8 < a.c) {
            #pragma unroll 1
            for (int pass = 0; pass < kDwTH / 4; ++pass) {
                const int ly = yr + 4 * pass, oy = ty0 + ly;
                const int lx0 = xblk * kDwOX, ox0 = tx0 + lx0;
                if (oy >= a.h || ox0 >= a.w) continue;
                float acc[kDwOX][8];
                {
                    const float4 b0 = *reinterpret_cast<const float4*>(s_b + g * 8), b1 = *reinterpret_cast<const float4*>(s_b + g * 8 + 4);
                    #pragma unroll
                    for (int j = 0; j < kDwOX; ++j) {
                        acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
                        acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
                    }
                }
                #pragma unroll
                for (int ky = 0; ky < K; ++ky) {
                    float wk[K][8];
                    #pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * 64 + g * 8);
                        const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * 64 + g * 8 + 4);
                        wk[kx][0] = w0.x; wk[kx][1] = w0.y; wk[kx][2] = w0.z; wk[kx][3] = w0.w;
                        wk[kx][4] = w1.x; wk[kx][5] = w1.y; wk[kx][6] = w1.z; wk[kx][7] = w1.w;
                    }
                    const uint4* row = tile + ((size_t)(ly + ky) * IW + lx0) * 8 + g;
                    #pragma unroll
                    for (int xi = 0; xi < kDwOX + K - 1; ++xi) {
                        const uint4 v = row[xi * 8];
                        const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y), bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
                        #pragma unroll
                        for (int j = 0; j < kDwOX; ++j) {
                            const int kx = xi - j;
                            if (kx >= 0 && kx < K) {
                                #pragma unroll
                                for (int q = 0; q < 8; ++q) acc[j][q] = fmaf(f[q], wk[kx][q], acc[j][q]);
                            }
                        }
                    }
                }
                __nv_bfloat16* yrow = a.y + ((size_t)(b * a.h + oy) * a.w + ox0) * a.yps + c0 + g * 8;
                #pragma unroll
                for (int j = 0; j < kDwOX; ++j) {
                    if (ox0 + j < a.w) {
                        uint4 o;
                        o.x = pack_bf16x2(silu_tanh(acc[j][0]), silu_tanh(acc[j][1]));
                        o.y = pack_bf16x2(silu_tanh(acc[j][2]), silu_tanh(acc[j][3]));
                        o.z = pack_bf16x2(silu_tanh(acc[j][4]), silu_tanh(acc[j][5]));
                        o.w = pack_bf16x2(silu_tanh(acc[j][6]), silu_tanh(acc[j][7]));
                        *reinterpret_cast<uint4*>(yrow + (size_t)j * a.yps) = o;
                    }
                }
            }
        }
        __syncthreads();                                  // the buffer is refilled by the prefetch of the next iteration
    }
}

template <int K>
int launch_dwconv(const __nv_bfloat16* xs, long long xps, int batch, int h, int w, int c, const float* weight, const float* bias,
                  __nv_bfloat16* ys, long long yps, cudaStream_t st) {
    constexpr int kDwTH = DwTile<K>::kTH;
    constexpr int IW = kDwTW + K - 1, IH = kDwTH + K - 1;
    const size_t smem = (size_t)2 * IH * IW * 8 * 16 + (size_t)(K * K * 64 + 64) * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(dwconv_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "dwconv: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    DwArgs a;
    a.x = xs; a.xps = xps; a.batch = batch; a.h = h; a.w = w; a.c = c; a.wgt = weight; a.bias = bias; a.y = ys; a.yps = yps;
    a.tiles_x = ceil_div(w, kDwTW); a.tiles_y = ceil_div(h, kDwTH); a.cblocks = ceil_div(c, 64);
    const long long total = (long long)a.tiles_x * a.tiles_y * batch * a.cblocks;
    if (total > 0x7fffffffLL) return fail(YMS_E_UNSUPPORTED, "dwconv: too many tiles");
    a.total = (int)total;
    const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
    const int grid = a.total < kNumSMs * ctas_per_sm ? a.total : kNumSMs * ctas_per_sm;
    dwconv_kernel<K><<<grid, kDwThreads, smem, st>>>(a);
    return check_launch("dwconv_kernel");
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
    if (c_out <= 0 || (c_out % 16) != 0 || c_out > 256) return fail(YMS_E_UNSUPPORTED, "stem: c_out must be a multiple of 16 (<= 256)");
    if (!x || !weight || !bias || !y || !aligned16(y) || (y_ps % 8) != 0) return fail(YMS_E_ARG, "stem: bad pointers/strides");
    if (c_out <= 128 && !getenv("YMS_STEM_LEGACY"))
        return yms_stem_tc_launch(x, nullptr, nullptr, nullptr, batch, in_h, in_w, c_out, weight, bias, y, y_ps, (cudaStream_t)stream);
    dim3 grid(ceil_div(in_w / 2, 128), in_h / 2, batch);
    size_t smem = (size_t)28 * c_out * sizeof(float);
    stem_conv_kernel<16><<<grid, 128, smem, (cudaStream_t)stream>>>(x, in_h, in_w, c_out, weight, bias,
                                                                    reinterpret_cast<__nv_bfloat16*>(y), y_ps);
    return check_launch("stem_conv_kernel");
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
        dim3 grid(c / 8, batch);
        const size_t smem = (size_t)2 * h * w * sizeof(uint4);
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(sppf_pool_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kPoolMaxHW * (int)sizeof(uint4));
            if (e != cudaSuccess) return fail((int)e, "sppf: smem attribute");
            attr_set = true;
        }
        sppf_pool_smem_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<__nv_bfloat16*>(buf), ps, h, w, c);
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

extern "C" int yms_dwconv(const void* x, int64_t xps, int batch, int h, int w, int channels, int ksize,
                          const float* weight, const float* bias, void* y, int64_t yps, void* stream) {
    if (batch <= 0 || h <= 0 || w <= 0 || channels <= 0 || (channels % 8) != 0) return fail(YMS_E_ARG, "dwconv: bad sizes (c % 8 == 0)");
    if (!x || !y || !weight || !bias || !aligned16(x) || !aligned16(y) || (xps % 8) != 0 || (yps % 8) != 0)
        return fail(YMS_E_ARG, "dwconv: bad pointers/strides");
    auto xs = reinterpret_cast<const __nv_bfloat16*>(x);
    auto ys = reinterpret_cast<__nv_bfloat16*>(y);
    cudaStream_t st = (cudaStream_t)stream;
    switch (ksize) {
        case 3: return launch_dwconv<3>(xs, xps, batch, h, w, channels, weight, bias, ys, yps, st);
        case 5: return launch_dwconv<5>(xs, xps, batch, h, w, channels, weight, bias, ys, yps, st);
        case 7: return launch_dwconv<7>(xs, xps, batch, h, w, channels, weight, bias, ys, yps, st);
        case 9: return launch_dwconv<9>(xs, xps, batch, h, w, channels, weight, bias, ys, yps, st);
        default: return fail(YMS_E_UNSUPPORTED, "dwconv: ksize must be 3, 5, 7 or 9");
    }
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
