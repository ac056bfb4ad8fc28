// Stem convolution backbone.conv0 (yolov8/model/yolov8_backbone.py:39; Conv = conv 3x3 stride 2
// + BN + SiLU, yolov8/model/components.py:69-77) on the tensor cores.
//
// Input is the caller's NCHW fp32 image, output NHWC bf16.  K = 3*3*3 = 27 is padded to 32:
//   * 8 producer warps gather the 27 taps of one output pixel per thread straight from the NCHW
//     image (next tile's loads are issued before the current tile is converted, so ~2 x 27 loads per
//     thread stay in flight), convert to bf16 and write the pixel's 64-byte K row into a
//     NON-swizzled K-major UMMA tile (8x16B core matrices: LBO = 128 B along K, SBO = 512 B along M);
//   * one elected thread issues 2 tcgen05.mma (M=128, N=c_out, K=16) per 128-pixel tile against the
//     weight tile that stays resident in shared memory, accumulating in TMEM (2 stages);
//   * 8 epilogue warps: tcgen05.ld -> +bias -> SiLU -> bf16 -> swizzled staging -> TMA store.
// HBM-bound: 12 B in + 2*c_out B out per output pixel.
#include "conv_plan.h"

#include <string.h>

namespace yms {
namespace {

using namespace tc;

constexpr int kProdWarps = 8;
constexpr int kStemThreads = kProdWarps * 32 + 32 + kEpiThreads;     // 544
constexpr int kStages = 4;
constexpr int kATile = 128 * 64;          // 8 KB: 128 pixels x 32 bf16
constexpr int kStageOutS = 16384;

struct StemParams {
    const float* x;            // NCHW fp32 image (kU8 == false)
    const unsigned char* xu8;  // NHWC uint8 image (kU8 == true): (v * scale[c] + shift[c]) == (v/255 - mean)/std
    float scale[3], shift[3];
    int batch, in_h, in_w, out_h, out_w, c_out;
    long long m_total; int total_tiles;
    const float* weight;     // f32 [c_out][27], BN folded
    const float* bias;       // f32 [c_out]
};

// no-swizzle K-major descriptor: start>>4 | LBO(128 B)>>4 <<16 | SBO(512 B)>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | (8ull << 16) | (32ull << 32) | (1ull << 46);
}

__device__ __forceinline__ void load_taps(const StemParams& p, long long pix, float (&v)[27]) {
    if (pix >= p.m_total) {
        #pragma unroll
        for (int i = 0; i < 27; ++i) v[i] = 0.f;
        return;
    }
    const int ox = (int)(pix % p.out_w);
    const long long t = pix / p.out_w;
    const int oy = (int)(t % p.out_h);
    const int b = (int)(t / p.out_h);
    const float* xb = p.x + (size_t)b * 3 * p.in_h * p.in_w;
    const int ix0 = 2 * ox - 1, iy0 = 2 * oy - 1;
    #pragma unroll
    for (int ci = 0; ci < 3; ++ci)
        #pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = iy0 + ky;
            const float* row = xb + ((size_t)ci * p.in_h + iy) * p.in_w + ix0;
            const bool yok = iy >= 0;
            v[ci * 9 + ky * 3 + 0] = (yok && ix0 >= 0) ? __ldg(row) : 0.f;
            v[ci * 9 + ky * 3 + 1] = yok ? __ldg(row + 1) : 0.f;
            v[ci * 9 + ky * 3 + 2] = yok ? __ldg(row + 2) : 0.f;
        }
}

// uint8 HWC input with the reference's ToTensor + Normalize (yolov8/tools/test.py:114-119) fused in:
// the 9 bytes of three horizontally adjacent RGB pixels are contiguous.  Padding taps are 0 AFTER
// normalisation (the reference pads the normalised tensor).
__device__ __forceinline__ void load_taps_u8(const StemParams& p, long long pix, float (&v)[27]) {
    #pragma unroll
    for (int i = 0; i < 27; ++i) v[i] = 0.f;
    if (pix >= p.m_total) return;
    const int ox = (int)(pix % p.out_w);
    const long long t = pix / p.out_w;
    const int oy = (int)(t % p.out_h);
    const int b = (int)(t / p.out_h);
    const int ix0 = 2 * ox - 1, iy0 = 2 * oy - 1;
    #pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = iy0 + ky;
        if (iy < 0) continue;
        const unsigned char* row = p.xu8 + (((long long)b * p.in_h + iy) * p.in_w + ix0) * 3;
        #pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            if (kx == 0 && ix0 < 0) continue;
            #pragma unroll
            for (int c = 0; c < 3; ++c)
                v[c * 9 + ky * 3 + kx] = fmaf((float)__ldg(row + kx * 3 + c), p.scale[c], p.shift[c]);
        }
    }
}

template <bool kU8>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ StemParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t smem_a = base;                                   // kStages x 8 KB
    const uint32_t smem_b = base + kStages * kATile;                // weights: c_out x 64 B (<= 8 KB)
    unsigned char* g_b = gbase + kStages * kATile;
    const uint32_t smem_out0 = smem_b + 8192;
    unsigned char* g_out0 = g_b + 8192;
    float* s_bias = reinterpret_cast<float*>(g_out0 + 2 * kStageOutS);          // 128 floats (0.5 * bias)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 128);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kStages + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pad = (p.c_out + 15) & ~15;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_y);
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 4); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == kProdWarps) tmem_alloc(smem_u32(tmem_slot), 256);
    // weights -> bf16 no-swizzle K-major tile: element (n, k) at (n/8)*512 + (k/8)*128 + (n%8)*16 + (k%8)*2
    for (int i = threadIdx.x; i < n_pad * 32; i += kStemThreads) {
        const int n = i >> 5, k = i & 31;
        const float w = (n < p.c_out && k < 27) ? p.weight[n * 27 + k] : 0.f;
        *reinterpret_cast<__nv_bfloat16*>(g_b + (n >> 3) * 512 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = __float2bfloat16(w);
    }
    for (int i = threadIdx.x; i < 128; i += kStemThreads) s_bias[i] = (i < p.c_out) ? 0.5f * p.bias[i] : 0.f;
    fence_proxy_async_smem();                                       // weight tile is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kProdWarps) {
        // ================= producers: im2col gather -> bf16 K rows =================
        const int grp = warp >> 2;                                  // 2 groups of 4 warps, alternate tiles
        const int r = (warp & 3) * 32 + lane;                       // row inside the tile
        float cur[27], nxt[27];
        int seq = grp;                                              // sequence number of this CTA's tiles
        long long t = (long long)blockIdx.x + (long long)grp * gridDim.x;
        if (t < p.total_tiles) { if (kU8) load_taps_u8(p, t * 128 + r, nxt); else load_taps(p, t * 128 + r, nxt); }
        for (; t < p.total_tiles; t += 2LL * gridDim.x, seq += 2) {
            #pragma unroll
            for (int i = 0; i < 27; ++i) cur[i] = nxt[i];
            const long long tn = t + 2LL * gridDim.x;
            if (tn < p.total_tiles) { if (kU8) load_taps_u8(p, tn * 128 + r, nxt); else load_taps(p, tn * 128 + r, nxt); }
            const int stage = seq % kStages;
            const uint32_t phase = (uint32_t)(seq / kStages) & 1u;
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t dst = smem_a + stage * kATile + (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 16u;
            #pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                uint32_t w[4];
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k0 = kc * 8 + q * 2;
                    const float a = (k0 < 27) ? cur[k0 < 27 ? k0 : 0] : 0.f;
                    const float c = (k0 + 1 < 27) ? cur[k0 + 1 < 27 ? k0 + 1 : 0] : 0.f;
                    w[q] = pack_bf16x2(a, c);
                }
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(dst + kc * 128u), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(stage));
        }
    } else if (warp == kProdWarps) {
        // ================= MMA issuer =================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t bdesc = make_nosw_desc(smem_b);
        int seq = 0, acc = 0; uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++seq) {
            const int stage = seq % kStages;
            const uint32_t phase = (uint32_t)(seq / kStages) & 1u;
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t adesc = make_nosw_desc(smem_a + stage * kATile);
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 128);
            if (elect_one()) {
                umma_bf16(d_tmem, adesc, bdesc, idesc, 0u);
                umma_bf16(d_tmem, adesc + 16ull, bdesc + 16ull, idesc, 1u);     // +256 B: next 16 K elements
                umma_commit(empty_bar(stage));
                umma_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - kProdWarps - 1;
        const int quad = warp & 3;
        const int half = ew >> 2;
        const int row = quad * 32 + lane;
        const bool leader = (ew == 0 && lane == 0);
        int acc = 0; uint32_t acc_phase = 0; uint32_t chunk_ctr = 0;
        const int n_chunks = (n_pad + 63) >> 6;
        for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 128);
            for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                const int buf = chunk_ctr & 1u;
                const uint32_t s_out = smem_out0 + buf * kStageOutS;
                const int c0 = ch * 64 + half * 32;
                const bool active = c0 < n_pad;
                if (leader) tma_store_wait_read<1>();
                epi_bar_sync();
                uint32_t v[32];
                if (active) { tmem_ld32(t_row + (uint32_t)c0, v); tmem_ld_wait(); }
                if (ch == n_chunks - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (active) {
                    float f[32];
                    const float4* bq = reinterpret_cast<const float4*>(s_bias + c0);
                    #pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 hb = bq[j];
                        f[4 * j + 0] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 0]), 0.5f, hb.x));
                        f[4 * j + 1] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 1]), 0.5f, hb.y));
                        f[4 * j + 2] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 2]), 0.5f, hb.z));
                        f[4 * j + 3] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 3]), 0.5f, hb.w));
                    }
                    const uint32_t line = s_out + (uint32_t)row * 128u;
                    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t addr = line + (((uint32_t)(half * 4 + q) ^ (uint32_t)(row & 7)) << 4);
                        const uint32_t o0 = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                        const uint32_t o1 = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                        const uint32_t o2 = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                        const uint32_t o3 = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                    }
                }
                fence_proxy_async_smem();
                epi_bar_sync();
                if (leader) {
                    tma_store_4d(&tm_y, s_out, ch * 64, (int)(t * 128), 0, 0);
                    tma_store_commit();
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (leader) tma_store_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kProdWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace
}  // namespace yms

using namespace yms;

// declared in glue.cu's dispatcher.  xu8 != nullptr selects the uint8-HWC + normalisation path.
int yms_stem_tc_launch(const float* x, const unsigned char* xu8, const float* mean, const float* stdv, int batch, int in_h, int in_w,
                       int c_out, const float* weight, const float* bias, void* y, int64_t y_ps, cudaStream_t stream) {
    static thread_local struct Cache { const void* y; int64_t ps; int b, h, w, c; CUtensorMap map; bool ok; } cache = {};
    const int out_h = in_h / 2, out_w = in_w / 2;
    if (!(cache.ok && cache.y == y && cache.ps == y_ps && cache.b == batch && cache.h == in_h && cache.w == in_w && cache.c == c_out)) {
        int rc = encode_act(&cache.map, y, c_out, y_ps, batch, out_h, out_w, /*flat=*/true, 128, 1, 1, "stem y");
        if (rc) return rc;
        cache.y = y; cache.ps = y_ps; cache.b = batch; cache.h = in_h; cache.w = in_w; cache.c = c_out; cache.ok = true;
    }
    StemParams p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.xu8 = xu8;
    if (xu8) for (int c = 0; c < 3; ++c) { p.scale[c] = 1.0f / (255.0f * stdv[c]); p.shift[c] = -mean[c] / stdv[c]; }
    p.batch = batch; p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.c_out = c_out;
    p.m_total = (long long)batch * out_h * out_w;
    p.total_tiles = (int)((p.m_total + 127) / 128);
    p.weight = weight; p.bias = bias;
    const size_t smem = 1024 + kStages * kATile + 8192 + 2 * kStageOutS + 128 * 4 + (2 * kStages + 4) * 8 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "stem: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
    if (xu8) stem_tc_kernel<true><<<grid, kStemThreads, smem, stream>>>(cache.map, p);
    else stem_tc_kernel<false><<<grid, kStemThreads, smem, stream>>>(cache.map, p);
    return check_launch("stem_tc_kernel");
}
