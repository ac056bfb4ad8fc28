// 3x3 stride-1 convolution (+ folded BN bias + SiLU + residual) as a tcgen05 implicit GEMM whose
// nine taps read ONE shared-memory halo tile through shifted UMMA descriptors.
//
// Replaces the 3x3 Conv units of Bottleneck.forward (yolov8/model/components.py:80-93) and of the
// head branches (yolov8/model/yolov8_head.py:84-85,99-100) of the reference.
//
// Why a second kernel: the generic kernel (conv_gemm.cu) fetches a shifted 16 KB activation tile
// and a weight tile from L2 for each of the 9 taps, i.e. ~10x more L2->SM traffic than the layer's
// HBM footprint; ncu showed those layers pinned at the L2->SM fabric rate (~8.5 TB/s), not at HBM.
// Here, per 64-channel block:
//   * the (th+2) x 10 pixel halo of an 8 x th (th <= 16) output sub-tile is loaded ONCE by a single
//     4-D TMA box (zero-filled outside the image = the conv padding); in shared memory it is
//     (th+2)*10 lines of 128 B (64 bf16 channels), SWIZZLE_128B;
//   * tap (ky,kx) is the SAME tile seen through a descriptor whose start address is advanced by
//     (ky*10 + kx) lines and whose 8-row group stride (SBO) is 10 lines = 1280 B: the 8 pixels of an
//     output row are 8 consecutive lines, consecutive output rows are 10 lines apart.  The 128 B
//     swizzle is a function of the absolute shared-memory address for both TMA and UMMA, so the
//     shifted views stay consistent;
//   * weights of small layers (9 * c_in/64 * c_out * 128 B <= ~110 KB) stay RESIDENT in shared
//     memory for the whole persistent CTA; larger ones stream through their own mbarrier ring and
//     are shared by two adjacent sub-tiles (two accumulators) to halve their L2 traffic.
// Warp roles / epilogue are the same as conv_gemm.cu.
#include "conv_plan.h"

#include <stdlib.h>
#include <string.h>

namespace yms {
namespace {

using namespace tc;

constexpr int kThreads3 = 64 + kEpiThreads;
constexpr int kHaloPitch = 10;                         // lines (pixels) per halo row
constexpr int kHaloStageBytes = 23552;                 // 180 lines = 23040 B, rounded up to 1024
constexpr int kStageOut = 16384;
constexpr int kRing = 8;                               // max ring depth (barrier array size)
constexpr int kSmemLimit3 = 232448;

// barrier slots
constexpr int kBarAFull = 0, kBarAEmpty = kRing, kBarBFull = 2 * kRing, kBarBEmpty = 3 * kRing;
constexpr int kBarTFull = 4 * kRing, kBarTEmpty = 4 * kRing + 2, kBarRes = 4 * kRing + 4, kBarW = 4 * kRing + 6;
constexpr int kNumBars = 4 * kRing + 8;

__device__ __forceinline__ uint64_t make_a_desc(uint32_t addr, int desc_mode) {
    uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((kHaloPitch * 128) >> 4) << 32) |
                 (1ull << 46) | (2ull << 61);
    if (desc_mode == 1) d |= (uint64_t)((addr >> 7) & 7u) << 49;
    return d;
}

struct Item { int n_tile, img, sx, ty; };
__device__ __forceinline__ Item decode_item(const Conv3Params& p, int t) {
    Item it;
    it.n_tile = t % p.n_tiles;
    int m = t / p.n_tiles;
    it.sx = m % p.super_x; m /= p.super_x;
    it.ty = m % p.tiles_y;
    it.img = m / p.tiles_y;
    return it;
}

__global__ void __launch_bounds__(kThreads3, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_res,
               const __grid_constant__ Conv3Params p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
    const int b_tile_bytes = (p.block_n * 128 + 1023) & ~1023;
    const int a_region = p.a_stages * p.sub * kHaloStageBytes;
    const int b_region = (p.resident ? 9 * p.kb : p.b_stages) * b_tile_bytes;
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + a_region;
    const uint32_t smem_out0 = smem_b + b_region;
    unsigned char* g_out0 = gbase + a_region + b_region;
    float* s_bias = reinterpret_cast<float*>(g_out0 + 2 * kStageOut);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int slot) { return bar0 + 8u * slot; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_x); prefetch_tmap(&tm_w); prefetch_tmap(&tm_y);
        if (p.has_res) prefetch_tmap(&tm_res);
        for (int s = 0; s < kRing; ++s) {
            mbar_init(bar(kBarAFull + s), 1); mbar_init(bar(kBarAEmpty + s), 1);
            mbar_init(bar(kBarBFull + s), 1); mbar_init(bar(kBarBEmpty + s), 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(bar(kBarTFull + s), 1); mbar_init(bar(kBarTEmpty + s), kEpiWarps); mbar_init(bar(kBarRes + s), 1); }
        mbar_init(bar(kBarW), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int i = threadIdx.x; i < p.bias_pad; i += kThreads3) s_bias[i] = (i < p.c_out) ? (p.act ? 0.5f * p.bias[i] : p.bias[i]) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t out_bytes = (uint32_t)(8 * p.th) * 128u;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            if (p.resident) {
                mbar_expect_tx(bar(kBarW), (uint32_t)(9 * p.kb) * (uint32_t)(p.block_n * 128));
                for (int tap = 0; tap < 9; ++tap)
                    for (int cb = 0; cb < p.kb; ++cb)
                        tma_load_3d(smem_b + (tap * p.kb + cb) * b_tile_bytes, &tm_w, bar(kBarW), cb * kBlockK, 0, tap);
            }
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
                const Item it = decode_item(p, t);
                const int n0 = it.n_tile * p.block_n;
                for (int cb = 0; cb < p.kb; ++cb) {
                    mbar_wait(bar(kBarAEmpty + as), aph ^ 1u);
                    mbar_expect_tx(bar(kBarAFull + as), (uint32_t)p.sub * p.halo_bytes);
                    for (int s = 0; s < p.sub; ++s)
                        tma_load_4d(smem_a + (as * p.sub + s) * kHaloStageBytes, &tm_x, bar(kBarAFull + as),
                                    cb * kBlockK, (it.sx * p.sub + s) * 8 - 1, it.ty * p.th - 1, it.img);
                    if (++as == p.a_stages) { as = 0; aph ^= 1u; }
                    if (!p.resident) {
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(bar(kBarBEmpty + bs), bph ^ 1u);
                            mbar_expect_tx(bar(kBarBFull + bs), (uint32_t)(p.block_n * 128));
                            tma_load_3d(smem_b + bs * b_tile_bytes, &tm_w, bar(kBarBFull + bs), cb * kBlockK, n0, tap);
                            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The issue loop is the critical path for small N (an MMA lasts only N/2 cycles): all
        // descriptor arithmetic is warp-uniform (uniform datapath) and hoisted out of the elected
        // region; taps and k-steps are fully unrolled so the operands are base + immediate.
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t hi_a = (1ull << 16) | ((uint64_t)((kHaloPitch * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t hi_b = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t halo16 = kHaloStageBytes >> 4;
        const uint32_t btile16 = (uint32_t)b_tile_bytes >> 4;
        if (p.resident) { mbar_wait(bar(kBarW), 0u); tc_fence_after(); }
        int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
            mbar_wait(bar(kBarTEmpty + acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
            for (int cb = 0; cb < p.kb; ++cb) {
                mbar_wait(bar(kBarAFull + as), aph);
                tc_fence_after();
                const int cvalid = p.c_in - cb * kBlockK;
                const int ksteps = cvalid >= kBlockK ? 4 : ((cvalid + 15) >> 4);
                const uint32_t a16 = ((smem_a + (uint32_t)(as * p.sub) * kHaloStageBytes) & 0x3FFFFu) >> 4;
                const uint32_t first = (cb != 0) ? 1u : 0u;
                const bool last_cb = (cb == p.kb - 1);
                if (p.resident) {
                    const uint32_t b16 = ((smem_b + (uint32_t)cb * b_tile_bytes) & 0x3FFFFu) >> 4;
                    const uint32_t bstride16 = (uint32_t)p.kb * btile16;
                    if (elect_one()) {
                        #pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t toff = (uint32_t)((tap / 3) * kHaloPitch + (tap % 3)) * 8u;
                            #pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps)
                                    umma_bf16(d_tmem, hi_a | (uint64_t)(a16 + toff + 2 * k), hi_b | (uint64_t)(b16 + tap * bstride16 + 2 * k),
                                              idesc, (tap | k) ? 1u : first);
                            }
                        }
                        umma_commit(bar(kBarAEmpty + as));
                        if (last_cb) umma_commit(bar(kBarTFull + acc));
                    }
                    __syncwarp();
                } else {
                    #pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(bar(kBarBFull + bs), bph);
                        tc_fence_after();
                        const uint32_t b16 = ((smem_b + (uint32_t)bs * b_tile_bytes) & 0x3FFFFu) >> 4;
                        const uint32_t toff = (uint32_t)((tap / 3) * kHaloPitch + (tap % 3)) * 8u;
                        const uint32_t accf = tap ? 1u : first;
                        if (elect_one()) {
                            #pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps)
                                    umma_bf16(d_tmem, hi_a | (uint64_t)(a16 + toff + 2 * k), hi_b | (uint64_t)(b16 + 2 * k), idesc, k ? 1u : accf);
                            }
                            if (p.sub == 2) {
                                #pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (k < ksteps)
                                        umma_bf16(d_tmem + (uint32_t)p.block_n, hi_a | (uint64_t)(a16 + halo16 + toff + 2 * k),
                                                  hi_b | (uint64_t)(b16 + 2 * k), idesc, k ? 1u : accf);
                                }
                            }
                            umma_commit(bar(kBarBEmpty + bs));
                            if (tap == 8) {
                                umma_commit(bar(kBarAEmpty + as));
                                if (last_cb) umma_commit(bar(kBarTFull + acc));
                            }
                        }
                        __syncwarp();
                        if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
                    }
                }
                if (++as == p.a_stages) { as = 0; aph ^= 1u; }
            }
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ================= epilogue (warps 2..9) =================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const bool leader = (threadIdx.x == 64);
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t chunk_ctr = 0;
        uint32_t res_phase0 = 0u, res_phase1 = 0u;
        const int n_chunks = (p.block_n + 63) >> 6;
        for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
            const Item it = decode_item(p, t);
            const int n0 = it.n_tile * p.block_n;
            const int y0 = it.ty * p.th;
            mbar_wait(bar(kBarTFull + acc), acc_phase);
            tc_fence_after();
            bool released = false;
            for (int s = 0; s < p.sub; ++s) {
                const int x0 = (it.sx * p.sub + s) * 8;
                if (x0 >= p.out_w) continue;                         // sub-tile entirely outside the image
                const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + s * p.block_n);
                for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                    const int buf = chunk_ctr & 1u;
                    const uint32_t s_out = smem_out0 + buf * kStageOut;
                    const int cbase = ch * 64;
                    const int c0 = cbase + half * 32;
                    const bool active = c0 < p.block_n;
                    if (leader) tma_store_wait_read<1>();
                    epi_bar_sync();
                    if (p.has_res) {
                        if (leader) {
                            mbar_expect_tx(bar(kBarRes + buf), out_bytes);
                            tma_load_4d(s_out, &tm_res, bar(kBarRes + buf), n0 + cbase, x0, y0, it.img);
                        }
                        const uint32_t ph = buf ? res_phase1 : res_phase0;
                        mbar_wait(bar(kBarRes + buf), ph);
                        if (buf) res_phase1 ^= 1u; else res_phase0 ^= 1u;
                    }
                    uint32_t v[32];
                    if (active) {
                        tmem_ld32(t_row + (uint32_t)c0, v);
                        tmem_ld_wait();
                    }
                    if (s == p.sub - 1 && ch == n_chunks - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(kBarTEmpty + acc));
                        released = true;
                    }
                    if (active) {
                        float f[32];
                        const float4* bq = reinterpret_cast<const float4*>(s_bias + n0 + c0);
                        if (p.act) {
                            #pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 hb = bq[j];
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
                    fence_proxy_async_smem();
                    epi_bar_sync();
                    if (leader) {
                        tma_store_4d(&tm_y, s_out, n0 + cbase, x0, y0, it.img);
                        tma_store_commit();
                    }
                }
            }
            if (!released) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(kBarTEmpty + acc));
            }
            if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
        }
        if (leader) tma_store_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int conv3_plan_init(yms_conv_plan* pl, const yms_conv_params* q) {
    Conv3Params& k = pl->k3;
    memset(&k, 0, sizeof(k));
    pl->kind = 1;
    const int H = q->in_h, W = q->in_w;
    k.out_w = W; k.out_h = H; k.batch = q->batch;
    const int ny = ceil_div(H, 16);
    k.th = ceil_div(H, ny);
    k.tiles_x = ceil_div(W, 8);
    k.tiles_y = ceil_div(H, k.th);
    k.c_in = q->c_in; k.c_out = q->c_out; k.kb = ceil_div(q->c_in, kBlockK);
    if (q->c_out <= 256) {
        k.n_tiles = 1; k.block_n = ((q->c_out + 15) / 16) * 16;
    } else {
        int best_pad = 1 << 30;
        for (int bn = 64; bn <= 256; bn += 64) {
            int padded = ceil_div(q->c_out, bn) * bn;
            if (padded <= best_pad) { best_pad = padded; k.block_n = bn; }
        }
        k.n_tiles = ceil_div(q->c_out, k.block_n);
    }
    k.act = q->act ? 1 : 0; k.has_res = q->residual ? 1 : 0;
    k.bias_pad = k.n_tiles * k.block_n + 64;
    k.bias = q->bias;
    k.halo_bytes = (uint32_t)(kHaloPitch * (k.th + 2) * 128);
    const char* dm = getenv("YMS_CONV3_DESC");
    k.desc_mode = dm ? atoi(dm) : 0;

    const int b_tile = (k.block_n * 128 + 1023) & ~1023;
    const int fixed = 2 * kStageOut + k.bias_pad * 4 + kNumBars * 8 + 16 + 1024;
    const int resident_bytes = 9 * k.kb * b_tile;
    const char* force_stream = getenv("YMS_CONV3_STREAM");
    k.resident = (k.n_tiles == 1 && kSmemLimit3 - fixed - resident_bytes >= 2 * kHaloStageBytes && !force_stream) ? 1 : 0;
    if (k.resident) {
        k.sub = 1; k.b_stages = 0;
        k.a_stages = (kSmemLimit3 - fixed - resident_bytes) / kHaloStageBytes;
    } else {
        k.sub = (k.tiles_x >= 2 && 2 * k.block_n <= 512) ? 2 : 1;
        k.b_stages = 4;
        for (;;) {
            k.a_stages = (kSmemLimit3 - fixed - k.b_stages * b_tile) / (k.sub * kHaloStageBytes);
            if (k.a_stages >= 2 || k.b_stages == 2) break;
            --k.b_stages;
        }
        if (k.a_stages < 2 && k.sub == 2) {
            k.sub = 1; k.b_stages = 4;
            k.a_stages = (kSmemLimit3 - fixed - k.b_stages * b_tile) / kHaloStageBytes;
        }
    }
    if (k.a_stages > kRing) k.a_stages = kRing;
    if (k.a_stages < 2) return fail(YMS_E_UNSUPPORTED, "conv3x3: tile does not fit in shared memory");
    k.acc_stages = (k.sub * k.block_n <= 256) ? 2 : 1;
    k.super_x = ceil_div(k.tiles_x, k.sub);
    k.total_items = k.super_x * k.tiles_y * k.batch * k.n_tiles;
    pl->grid = k.total_items < kNumSMs ? k.total_items : kNumSMs;
    pl->smem = (size_t)k.a_stages * k.sub * kHaloStageBytes + (size_t)(k.resident ? resident_bytes : k.b_stages * b_tile) + fixed;

    int rc;
    if ((rc = encode_act(&pl->tm_x, q->x, q->c_in, q->x_pixel_stride, q->batch, H, W, false, kHaloPitch, k.th + 2, 1, "x(halo)"))) return rc;
    {
        uint64_t dims[3] = {(uint64_t)q->c_in, (uint64_t)q->c_out, 9};
        uint64_t strides[2] = {(uint64_t)q->c_in * 2, (uint64_t)q->c_in * 2 * (uint64_t)q->c_out};
        uint32_t box[3] = {kBlockK, (uint32_t)k.block_n, 1};
        uint32_t es[3] = {1, 1, 1};
        if ((rc = encode_map(&pl->tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->weight, dims, strides, box, es, "w"))) return rc;
    }
    if ((rc = encode_act(&pl->tm_y, q->y, q->c_out, q->y_pixel_stride, q->batch, H, W, false, 8, k.th, 1, "y"))) return rc;
    if (q->residual) {
        if ((rc = encode_act(&pl->tm_res, q->residual, q->c_out, q->res_pixel_stride, q->batch, H, W, false, 8, k.th, 1, "res"))) return rc;
    } else pl->tm_res = pl->tm_y;
    pl->tm_x2 = pl->tm_x;

    const double m = (double)q->batch * H * W;
    pl->flops = 2.0 * m * q->c_out * (double)q->c_in * 9.0;
    pl->bytes = 2.0 * m * q->c_in + 2.0 * m * q->c_out + 2.0 * 9.0 * q->c_out * q->c_in + (q->residual ? 2.0 * m * q->c_out : 0.0);

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        if (e != cudaSuccess) return fail((int)e, "conv3x3: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    return 0;
}

int conv3_plan_run(const yms_conv_plan* pl, cudaStream_t stream) {
    conv3x3_kernel<<<pl->grid, kThreads3, pl->smem, stream>>>(pl->tm_x, pl->tm_w, pl->tm_y, pl->tm_res, pl->k3);
    return check_launch("conv3x3_kernel");
}

}  // namespace yms
