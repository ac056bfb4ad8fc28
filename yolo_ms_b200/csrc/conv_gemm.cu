// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA, folded-BN bias + SiLU (+ residual) fused into the epilogue.
//
// Replaces Conv.forward (yolov8/model/components.py:69-77), the Bottleneck residual (:87-93),
// and -- because inputs/outputs are channel slices addressed through tensor maps -- the
// torch.cat / slicing traffic of C2f.forward (:108-122) and Neck.forward (yolov8_neck.py:76-92).
//
// GEMM view:  D[M = output pixels, N = c_out] = sum over taps (ky,kx) and 64-channel blocks of
//             A[M, 64] (shifted NHWC activation tile)  x  B[N, 64]^T (weights of that tap).
//   * A tile: ONE 4-D TMA box (64 ch, TW, TH, 1 image) per (tap, channel block); the box origin
//     is shifted by the tap offset, out-of-image pixels and channels >= c_in are zero-filled
//     by the TMA unit (that is the conv padding).  Stride-2 convs use the tensor map's
//     traversal stride (elementStrides = 2) so the box still lands as TW x TH dense rows.
//     1x1 convs use a flat view (M = B*H*W, TW = 128, TH = 1).
//   * smem tiles are 128-byte rows, SWIZZLE_128B, K-major: exactly the canonical UMMA layout.
//   * persistent CTAs (one per SM), static round-robin tile scheduler, warp-specialised:
//       warp 0     TMA producer            (smem full/empty mbarrier ring, 3..8 stages)
//       warp 1     tcgen05.mma issuer      (one elected lane), owns the TMEM allocation
//       warps 2-9  epilogue (2 warps per TMEM lane quadrant, 32 columns each per 64-column chunk):
//                  tcgen05.ld -> +bias -> SiLU -> (+residual) -> bf16 -> swizzled smem -> TMA store
//                  (or direct fp32 stores for the head's raw logits)
//     two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include "conv_plan.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <new>

namespace yms {
namespace {

constexpr int kBlockM = 128;
constexpr int kMaxStages = 8;
constexpr int kATileBytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kStageOutBytes = kBlockM * 128;           // 16 KB epilogue staging (64 bf16 ch per row)
using namespace tc;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;                // TMEM columns between the two accumulator stages
constexpr int kSmemLimit = 232448;             // 227 KB

struct TileCoord { int n_tile, img, x0, y0; };
__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int t) {
    TileCoord c;
    c.n_tile = t % p.n_tiles;
    int m = t / p.n_tiles;
    int tx = m % p.tiles_x; m /= p.tiles_x;
    int ty = m % p.tiles_y;
    c.img = m / p.tiles_y;
    c.x0 = tx * p.tw; c.y0 = ty * p.th;
    return c;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x2,
                 const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
                 const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ ConvKernelParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // carve-up (1024-byte aligned for SWIZZLE_128B)
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
    const int b_tile_bytes = p.block_n * 128;
    const int b_tile_pad = (b_tile_bytes + 1023) & ~1023;
    // streaming: stage = A tile + B tile; resident: stage = A tile only, all B tiles after the ring
    const int stage_bytes = kATileBytes + (p.resident ? 0 : b_tile_pad);
    const int kb_total_res = p.taps * (p.kb1 + p.kb2);
    const int ring_bytes = p.num_stages * stage_bytes + (p.resident ? kb_total_res * b_tile_pad : 0);
    const uint32_t smem_a0 = base;                                   // stage s: A at base + s*stage_bytes, B after A
    const uint32_t smem_bres = base + p.num_stages * stage_bytes;     // resident weight tiles
    const uint32_t smem_out0 = base + ring_bytes;                     // 2 x 16 KB staging
    unsigned char* g_out0 = gbase + ring_bytes;
    float* s_bias = reinterpret_cast<float*>(g_out0 + 2 * kStageOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 2 + s); };
    auto res_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 4 + s); };
    auto w_bar = [&]() { return bar0 + 8u * (2 * kMaxStages + 6); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 7);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_x); prefetch_tmap(&tm_w);
        if (p.kb2) prefetch_tmap(&tm_x2);
        if (!p.out_f32) prefetch_tmap(&tm_y);
        if (p.has_res) prefetch_tmap(&tm_res);
        for (int s = 0; s < p.num_stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); mbar_init(res_bar(s), 1); }
        mbar_init(w_bar(), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    for (int i = threadIdx.x; i < p.bias_pad; i += kThreads) s_bias[i] = (i < p.c_out) ? (p.act ? 0.5f * p.bias[i] : p.bias[i]) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kb_per_tap = p.kb1 + p.kb2;
    const int num_kb = p.taps * kb_per_tap;
    const int pad = p.ksize >> 1;
    const uint32_t a_bytes = (uint32_t)(p.tw * p.th) * 128u;
    const uint32_t stage_tx = a_bytes + (p.resident ? 0u : (uint32_t)b_tile_bytes);

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            if (p.resident) {                                   // whole weight set once per persistent CTA
                mbar_expect_tx(w_bar(), (uint32_t)kb_total_res * (uint32_t)b_tile_bytes);
                for (int tap = 0; tap < p.taps; ++tap)
                    for (int kb = 0; kb < kb_per_tap; ++kb)
                        tma_load_3d(smem_bres + (tap * kb_per_tap + kb) * b_tile_pad, &tm_w, w_bar(),
                                    kb < p.kb1 ? kb * kBlockK : p.c_in1 + (kb - p.kb1) * kBlockK, 0, tap);
            }
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t);
                const int n0 = tc.n_tile * p.block_n;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
                    const int xin = tc.x0 * p.stride + kx - pad;
                    const int yin = tc.y0 * p.stride + ky - pad;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t sa = smem_a0 + stage * stage_bytes;
                        const uint32_t sb = sa + kATileBytes;
                        mbar_expect_tx(full_bar(stage), stage_tx);
                        if (kb < p.kb1) {
                            tma_load_4d(sa, &tm_x, full_bar(stage), kb * kBlockK, xin, yin, tc.img);
                            if (!p.resident) tma_load_3d(sb, &tm_w, full_bar(stage), kb * kBlockK, n0, tap);
                        } else {
                            tma_load_4d(sa, &tm_x2, full_bar(stage), (kb - p.kb1) * kBlockK, xin, yin, tc.img);
                            if (!p.resident) tma_load_3d(sb, &tm_w, full_bar(stage), p.c_in1 + (kb - p.kb1) * kBlockK, n0, tap);
                        }
                        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
        if (p.resident) { mbar_wait(w_bar(), 0u); tc_fence_after(); }
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
            int kbi = 0;
            for (int tap = 0; tap < p.taps; ++tap) {
                for (int kb = 0; kb < kb_per_tap; ++kb, ++kbi) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    // warp-uniform descriptor arithmetic outside the elected region (uniform datapath)
                    const int cvalid = (kb < p.kb1) ? (p.c_in1 - kb * kBlockK) : (p.c_in2 - (kb - p.kb1) * kBlockK);
                    const int ksteps = cvalid >= kBlockK ? 4 : ((cvalid + 15) >> 4);
                    const uint32_t sa = smem_a0 + stage * stage_bytes;
                    const uint64_t adesc = make_sw128_desc(sa);
                    const uint64_t bdesc = make_sw128_desc(p.resident ? smem_bres + (uint32_t)kbi * b_tile_pad : sa + kATileBytes);
                    const uint32_t first = kbi ? 1u : 0u;
                    const bool last = (kbi == num_kb - 1);
                    if (elect_one()) {
                        #pragma unroll
                        for (int k = 0; k < 4; ++k) {              // +32 B (16 bf16) along K inside the swizzle atom
                            if (k < ksteps) umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : first);
                        }
                        umma_commit(empty_bar(stage));             // smem slot free once these MMAs retire
                        if (last) umma_commit(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ================= epilogue (warps 2..9) =================
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                  // which 32 columns of every 64-column chunk
        const int row = quad * 32 + lane;                  // tile row == accumulator lane
        const bool leader = (threadIdx.x == 64);
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t chunk_ctr = 0;
        uint32_t res_phase0 = 0u, res_phase1 = 0u;
        const int n_chunks = (p.block_n + 63) >> 6;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(p, t);
            const int n0 = tc.n_tile * p.block_n;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccStride);
            // fp32 path: this thread's output pixel
            float* yrow = nullptr;
            if (p.out_f32) {
                const int lty = row / p.tw, ltx = row - lty * p.tw;
                if ((row < p.tw * p.th) && (tc.x0 + ltx < p.out_w) && (tc.y0 + lty < p.out_h))
                    yrow = p.y_f32 + ((size_t)((size_t)tc.img * p.out_h + tc.y0 + lty) * p.out_w + tc.x0 + ltx) * p.y_ps;
            }
            for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                const int buf = chunk_ctr & 1u;
                const uint32_t s_out = smem_out0 + buf * kStageOutBytes;
                const int cbase = ch * 64;                           // column inside the N tile
                const int c0 = cbase + half * 32;
                const bool active = c0 < p.block_n;                  // warp-uniform
                if (!p.out_f32) {
                    if (leader) tma_store_wait_read<1>();           // staging buffer `buf` no longer being read
                    epi_bar_sync();
                    if (p.has_res) {
                        if (leader) {
                            mbar_expect_tx(res_bar(buf), a_bytes);
                            tma_load_4d(s_out, &tm_res, res_bar(buf), n0 + cbase, tc.x0, tc.y0, tc.img);
                        }
                        const uint32_t ph = buf ? res_phase1 : res_phase0;
                        mbar_wait(res_bar(buf), ph);
                        if (buf) res_phase1 ^= 1u; else res_phase0 ^= 1u;
                    }
                }
                uint32_t v[32];
                if (active) {
                    tmem_ld32(t_row + (uint32_t)c0, v);
                    tmem_ld_wait();
                }
                if (ch == n_chunks - 1) {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
                if (active) {
                    float f[32];
                    const float4* bq = reinterpret_cast<const float4*>(s_bias + n0 + c0);
                    if (p.act) {
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 hb = bq[j];                 // 0.5 * bias
                            f[4 * j + 0] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 0]), 0.5f, hb.x));
                            f[4 * j + 1] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 1]), 0.5f, hb.y));
                            f[4 * j + 2] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 2]), 0.5f, hb.z));
                            f[4 * j + 3] = silu_from_half(fmaf(__uint_as_float(v[4 * j + 3]), 0.5f, hb.w));
                        }
                    } else {
                        #pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = bq[j];
                            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4.x;
                            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
                            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
                            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
                        }
                    }
                    if (p.out_f32) {
                        if (yrow) {
                            #pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const int col = n0 + c0 + j;
                                if (col + 3 < p.c_out) {
                                    *reinterpret_cast<float4*>(yrow + col) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                                } else {
                                    for (int q = 0; q < 4; ++q) if (col + q < p.c_out) yrow[col + q] = f[j + q];
                                }
                            }
                        }
                    } else {
                        // 32 columns = 64 B = 4 x 16 B chunks of this row's 128 B swizzled line
                        const uint32_t line = s_out + (uint32_t)row * 128u;
                        #pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t chunk16 = (uint32_t)(half * 4 + q);
                            const uint32_t addr = line + ((chunk16 ^ (uint32_t)(row & 7)) << 4);
                            if (p.has_res) {
                                uint32_t r0, r1, r2, r3;
                                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
                                f[q * 8 + 0] += bf16_lo(r0); f[q * 8 + 1] += bf16_hi(r0);
                                f[q * 8 + 2] += bf16_lo(r1); f[q * 8 + 3] += bf16_hi(r1);
                                f[q * 8 + 4] += bf16_lo(r2); f[q * 8 + 5] += bf16_hi(r2);
                                f[q * 8 + 6] += bf16_lo(r3); f[q * 8 + 7] += bf16_hi(r3);
                            }
                            const uint32_t o0 = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                            const uint32_t o1 = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                            const uint32_t o2 = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                            const uint32_t o3 = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                        }
                    }
                }
                if (!p.out_f32) {
                    fence_proxy_async_smem();                        // generic-proxy writes -> async proxy (TMA)
                    epi_bar_sync();
                    if (leader) {
                        tma_store_4d(&tm_y, s_out, n0 + cbase, tc.x0, tc.y0, tc.img);
                        tma_store_commit();
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (leader && !p.out_f32) tma_store_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
}  // namespace

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

int encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* addr, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr, const char* what) {
    auto fn = get_encode();
    if (!fn) return fail(YMS_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(addr), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(YMS_E_DRIVER, "cuTensorMapEncodeTiled(%s) failed: %d (dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", what, (int)r,
                    (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                    (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 0;
}


// activation tensor map: dims (c, X, Y, N); `flat` folds all pixels into X.
int encode_act(CUtensorMap* m, const void* ptr, int c, int64_t ps, int batch, int h, int w, bool flat,
               int box_x, int box_y, int estride, const char* what) {
    uint64_t dims[4]; uint64_t strides[3]; uint32_t box[4]; uint32_t es[4] = {1, (uint32_t)estride, (uint32_t)estride, 1};
    dims[0] = (uint64_t)c;
    strides[0] = (uint64_t)ps * 2;
    if (flat) {
        dims[1] = (uint64_t)batch * h * w; dims[2] = 1; dims[3] = 1;
        strides[1] = strides[0] * dims[1]; strides[2] = strides[1];
    } else {
        dims[1] = (uint64_t)w; dims[2] = (uint64_t)h; dims[3] = (uint64_t)batch;
        strides[1] = strides[0] * (uint64_t)w; strides[2] = strides[1] * (uint64_t)h;
    }
    box[0] = kBlockK; box[1] = (uint32_t)box_x; box[2] = (uint32_t)box_y; box[3] = 1;
    return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, what);
}



}  // namespace yms

using namespace yms;

extern "C" int yms_conv_plan_create(const yms_conv_params* q, yms_conv_plan** out) {
    if (!q || !out) return fail(YMS_E_ARG, "conv: null argument");
    *out = nullptr;
    if (q->batch <= 0 || q->in_h <= 0 || q->in_w <= 0 || q->c_in <= 0 || q->c_out <= 0 || q->c_in2 < 0)
        return fail(YMS_E_ARG, "conv: bad sizes");
    if (!((q->ksize == 1 && q->stride == 1) || (q->ksize == 3 && (q->stride == 1 || q->stride == 2))))
        return fail(YMS_E_UNSUPPORTED, "conv: only 1x1/s1, 3x3/s1, 3x3/s2 are implemented");
    if (q->stride == 2 && ((q->in_h | q->in_w) & 1)) return fail(YMS_E_UNSUPPORTED, "conv: stride 2 needs even H, W");
    if ((q->c_in % 8) || (q->c_in2 % 8) || (q->c_out % 8)) return fail(YMS_E_UNSUPPORTED, "conv: channels must be multiples of 8");
    if (q->out_dtype != YMS_DTYPE_BF16 && q->out_dtype != YMS_DTYPE_F32) return fail(YMS_E_UNSUPPORTED, "conv: out dtype");
    if (q->out_dtype == YMS_DTYPE_F32 && q->residual) return fail(YMS_E_UNSUPPORTED, "conv: residual needs bf16 output");
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    if (!q->x || !q->y || !q->weight || !q->bias || !al16(q->x) || !al16(q->y) || !al16(q->weight))
        return fail(YMS_E_ARG, "conv: null or misaligned pointer");
    if ((q->x_pixel_stride % 8) || q->x_pixel_stride < q->c_in || q->y_pixel_stride < q->c_out)
        return fail(YMS_E_ARG, "conv: bad pixel stride");
    if (q->out_dtype == YMS_DTYPE_BF16 && (q->y_pixel_stride % 8)) return fail(YMS_E_ARG, "conv: y pixel stride % 8");
    if (q->out_dtype == YMS_DTYPE_F32 && (q->y_pixel_stride % 4)) return fail(YMS_E_ARG, "conv: y pixel stride % 4");
    if (q->c_in2 && (!q->x2 || !al16(q->x2) || (q->x2_pixel_stride % 8) || q->x2_pixel_stride < q->c_in2))
        return fail(YMS_E_ARG, "conv: bad second source");
    if (q->residual && (!al16(q->residual) || (q->res_pixel_stride % 8) || q->res_pixel_stride < q->c_out))
        return fail(YMS_E_ARG, "conv: bad residual");

    yms_conv_plan* pl = new (std::nothrow) yms_conv_plan();
    if (!pl) return fail(YMS_E_ARG, "conv: out of host memory");
    pl->kind = 0;
    if (q->ksize == 3 && q->stride == 1 && q->out_dtype == YMS_DTYPE_BF16 && q->c_in2 == 0 && !getenv("YMS_CONV3_LEGACY")) {
        int rc3 = conv3_plan_init(pl, q);
        if (rc3) { delete pl; return rc3; }
        *out = pl;
        return 0;
    }
    ConvKernelParams& kp = pl->kp;
    memset(&kp, 0, sizeof(kp));
    const int out_h = q->in_h / q->stride, out_w = q->in_w / q->stride;
    const bool flat = (q->ksize == 1);
    if (flat) {
        kp.tw = kBlockM; kp.th = 1;
        kp.out_w = q->batch * out_h * out_w; kp.out_h = 1; kp.batch = 1;
    } else {
        // pick the output tile (tw x th <= 128 pixels) that wastes the fewest MMA rows
        double best = -1.0; int btw = 1, bth = 1;
        for (int tw = 1; tw <= (out_w < 128 ? out_w : 128); ++tw) {
            int th = 128 / tw; if (th > out_h) th = out_h;
            if (th * q->stride > 256 || tw * q->stride > 256) continue;
            double eff = (double)out_w * out_h / ((double)ceil_div(out_w, tw) * ceil_div(out_h, th) * 128.0);
            if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > btw)) { best = eff; btw = tw; bth = th; }
        }
        kp.tw = btw; kp.th = bth; kp.out_w = out_w; kp.out_h = out_h; kp.batch = q->batch;
    }
    kp.tiles_x = ceil_div(kp.out_w, kp.tw);
    kp.tiles_y = ceil_div(kp.out_h, kp.th);
    if (q->c_out <= 256) {
        kp.n_tiles = 1;
        kp.block_n = ((q->c_out + 15) / 16) * 16;
    } else {
        // several N tiles: every tile must be whole 64-channel store boxes (a partial box would
        // spill into the next tile's channels); pick the width that pads c_out the least.
        int best_pad = 1 << 30;
        for (int bn = 64; bn <= 256; bn += 64) {
            int padded = ceil_div(q->c_out, bn) * bn;
            if (padded <= best_pad) { best_pad = padded; kp.block_n = bn; }
        }
        kp.n_tiles = ceil_div(q->c_out, kp.block_n);
    }
    kp.c_out = q->c_out;
    kp.c_in1 = q->c_in; kp.c_in2 = q->c_in2;
    kp.kb1 = ceil_div(q->c_in, kBlockK); kp.kb2 = ceil_div(q->c_in2, kBlockK);
    kp.ksize = q->ksize; kp.taps = q->ksize * q->ksize; kp.stride = q->stride;
    kp.act = q->act ? 1 : 0; kp.out_f32 = (q->out_dtype == YMS_DTYPE_F32); kp.has_res = q->residual ? 1 : 0;
    kp.total_tiles = kp.tiles_x * kp.tiles_y * kp.batch * kp.n_tiles;
    kp.bias_pad = kp.n_tiles * kp.block_n + 64;
    kp.bias = q->bias;
    kp.y_f32 = kp.out_f32 ? reinterpret_cast<float*>(q->y) : nullptr;
    kp.y_ps = q->y_pixel_stride;

    const int b_bytes = (kp.block_n * 128 + 1023) & ~1023;
    const int fixed = 2 * kStageOutBytes + kp.bias_pad * 4 + (2 * kMaxStages + 8) * 8 + 1024 /* alignment slack */;
    const int res_bytes = kp.taps * (kp.kb1 + kp.kb2) * b_bytes;
    // small weight sets stay resident for the whole persistent CTA (no per-tile re-fetch from L2)
    kp.resident = (kp.n_tiles == 1 && !getenv("YMS_CONV_STREAM") && kSmemLimit - fixed - res_bytes >= 4 * kATileBytes) ? 1 : 0;
    const int stage_bytes = kATileBytes + (kp.resident ? 0 : b_bytes);
    int stages = (kSmemLimit - fixed - (kp.resident ? res_bytes : 0)) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { delete pl; return fail(YMS_E_UNSUPPORTED, "conv: tile does not fit in shared memory"); }
    kp.num_stages = stages;
    pl->smem = (size_t)stages * stage_bytes + (kp.resident ? res_bytes : 0) + fixed;
    pl->grid = kp.total_tiles < kNumSMs ? kp.total_tiles : kNumSMs;

    int rc;
    const int K_total = q->c_in + q->c_in2;
    const int bx = kp.tw * q->stride, by = kp.th * q->stride;
    if ((rc = encode_act(&pl->tm_x, q->x, q->c_in, q->x_pixel_stride, q->batch, q->in_h, q->in_w, flat, bx, by, q->stride, "x"))) { delete pl; return rc; }
    if (q->c_in2) {
        if ((rc = encode_act(&pl->tm_x2, q->x2, q->c_in2, q->x2_pixel_stride, q->batch, q->in_h, q->in_w, flat, bx, by, q->stride, "x2"))) { delete pl; return rc; }
    } else pl->tm_x2 = pl->tm_x;
    {
        uint64_t dims[3] = {(uint64_t)K_total, (uint64_t)q->c_out, (uint64_t)kp.taps};
        uint64_t strides[2] = {(uint64_t)K_total * 2, (uint64_t)K_total * 2 * (uint64_t)q->c_out};
        uint32_t box[3] = {kBlockK, (uint32_t)kp.block_n, 1};
        uint32_t es[3] = {1, 1, 1};
        if ((rc = encode_map(&pl->tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->weight, dims, strides, box, es, "w"))) { delete pl; return rc; }
    }
    if (!kp.out_f32) {
        if ((rc = encode_act(&pl->tm_y, q->y, q->c_out, q->y_pixel_stride, q->batch, out_h, out_w, flat, kp.tw, kp.th, 1, "y"))) { delete pl; return rc; }
    } else pl->tm_y = pl->tm_x;
    if (q->residual) {
        if ((rc = encode_act(&pl->tm_res, q->residual, q->c_out, q->res_pixel_stride, q->batch, out_h, out_w, flat, kp.tw, kp.th, 1, "res"))) { delete pl; return rc; }
    } else pl->tm_res = pl->tm_x;

    const double m = (double)q->batch * out_h * out_w;
    pl->flops = 2.0 * m * q->c_out * (double)K_total * kp.taps;
    pl->bytes = 2.0 * (double)q->batch * q->in_h * q->in_w * K_total + (kp.out_f32 ? 4.0 : 2.0) * m * q->c_out +
                2.0 * (double)kp.taps * q->c_out * K_total + (q->residual ? 2.0 * m * q->c_out : 0.0);

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) { delete pl; return fail((int)e, "conv: smem attribute: %s", cudaGetErrorString(e)); }
        attr_set = true;
    }
    *out = pl;
    return 0;
}

extern "C" int yms_conv_plan_run(const yms_conv_plan* pl, void* stream) {
    if (!pl) return fail(YMS_E_ARG, "conv: null plan");
    if (pl->kind == 1) return conv3_plan_run(pl, (cudaStream_t)stream);
    conv_gemm_kernel<<<pl->grid, kThreads, pl->smem, (cudaStream_t)stream>>>(pl->tm_x, pl->tm_x2, pl->tm_w, pl->tm_y,
                                                                             pl->tm_res, pl->kp);
    return check_launch("conv_gemm_kernel");
}

extern "C" int yms_conv_plan_destroy(yms_conv_plan* pl) {
    delete pl;
    return 0;
}

extern "C" int yms_conv_plan_cost(const yms_conv_plan* pl, double* flops, double* bytes) {
    if (!pl) return fail(YMS_E_ARG, "conv: null plan");
    if (flops) *flops = pl->flops;
    if (bytes) *bytes = pl->bytes;
    return 0;
}
