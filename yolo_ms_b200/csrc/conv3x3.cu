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
//
// Stride-2 "pair-line" mode (s2pair; backbone.conv1 of the s model: 3x3/s2, c_in = 32, dense NHWC input).  The generic
// kernel needs 9 shifted 128-row TMA boxes per output tile there and is bound by the TMA row rate (~4 cycles per 128 B
// row, half of each row being zero fill).  A dense stride-2 input viewed as (2C = 64 channels, W/2 pixel pairs, 2 row
// parities, H/2, N) has 128-byte lines that hold BOTH pixels of a pair, so:
//   * per tile only the two parity planes are loaded: 2 boxes of 9 pairs x (th+1) rows (4x fewer rows);
//   * taps (ky,1),(ky,2) are ONE K = 64 block of the pair line at output x (weights [w(ky,1) | w(ky,2)]), tap (ky,0) is
//     the upper 32 channels of the pair line at x-1 (k-steps 2,3 against [0 | w(ky,0)]): 6 weight tiles ("pair-packed",
//     yms_conv_params.variant == 4) and 18 MMAs per tile, all on shifted descriptors of the two planes (row pitch 9).
#include "conv_plan.h"

#include <stdlib.h>
#include <string.h>

namespace yms {
namespace {

using namespace tc;

constexpr int kThreads3 = kConvThreads;
constexpr int kHaloPitch = 10;                         // lines (pixels) per halo row
constexpr int kRing = 8;                               // max ring depth (barrier array size)
constexpr int kSmemLimit3 = 232448;

// barrier slots
constexpr int kBarAFull = 0, kBarAEmpty = kRing, kBarBFull = 2 * kRing, kBarBEmpty = 3 * kRing;
constexpr int kBarTFull = 4 * kRing, kBarTEmpty = 4 * kRing + 4, kBarRes = 4 * kRing + 8, kBarW = 4 * kRing + 12;
constexpr int kNumBars = 4 * kRing + 14;

struct Item { int n_tile, img, sx, ty; };
__device__ __forceinline__ Item decode_item(const Conv3Params& p, int t) {
    Item it;
    uint32_t m = fast_div((uint32_t)t, p.mg_n_tiles);
    it.n_tile = t - (int)m * p.n_tiles;
    uint32_t q = fast_div(m, p.mg_super_x);
    it.sx = (int)(m - q * p.super_x);
    m = q;
    q = fast_div(m, p.mg_tiles_y);
    it.ty = (int)(m - q * p.tiles_y);
    it.img = (int)q;
    return it;
}

template <int kSub>
__global__ void __launch_bounds__(kThreads3, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_res,
               const __grid_constant__ Conv3Params p) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];      // SWIZZLE_128B tiles need 1024-byte alignment
    long long pw0 = 0, pw1 = 0, pw2 = 0, pw3 = 0;   // wait-cycle accumulators (dead code unless -DYMS_PROF)
    (void)pw0; (void)pw1; (void)pw2; (void)pw3;
    YMS_PROF_ONLY(const long long prof_t_entry = clock64(); long long* prof = p.prof ? p.prof + 16 * blockIdx.x : nullptr;)
    const uint32_t base = smem_u32(smem_dyn);
    unsigned char* gbase = smem_dyn;
    if (base & 1023u) __trap();
    const int b_tile_bytes = (p.block_n * 128 + 1023) & ~1023;
    const int a_region = ((p.a_stages * p.planes * p.halo_stage) + 1023) & ~1023;
    const int b_region = (p.resident ? p.wtiles * p.kb : p.b_stages) * b_tile_bytes;
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + a_region;
    const uint32_t smem_out0 = smem_b + b_region;
    unsigned char* g_out0 = gbase + a_region + b_region;
    float* s_bias = reinterpret_cast<float*>(g_out0 + kEpiGroups * kStageOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int slot) { return bar0 + 8u * slot; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const BiasRegs bias_regs = bias_fetch(p.bias, p.c_out);          // staged by the epilogue warps after the CTA-wide sync
    if (warp == 0) {                                  // barriers initialised lane-parallel: the prologue is paid by every launch
        if (lane == 0) {
            prefetch_tmap(&tm_x); prefetch_tmap(&tm_w); prefetch_tmap(&tm_y);
            if (p.has_res) prefetch_tmap(&tm_res);
        }
        for (int i = lane; i < kNumBars; i += 32) {
            const bool is_tempty = i >= kBarTEmpty && i < kBarTEmpty + 4;
            mbar_init(bar(i), is_tempty ? 4 * (kEpiGroups / p.acc_stages) : 1);
        }
        fence_barrier_init();
        __syncwarp();
        // weights are constants of the program: fetched before the CTA-wide sync (overlapping the TMEM allocation and the
        // bias staging) and BEFORE the grid dependency resolves, i.e. while the previous layer is still draining
        if (p.resident && elect_one()) {
            mbar_expect_tx(bar(kBarW), (uint32_t)(p.wtiles * p.kb) * (uint32_t)(p.block_n * 128));
            for (int tap = 0; tap < p.wtiles; ++tap)
                for (int cb = 0; cb < p.kb; ++cb)
                    tma_load_3d(smem_b + (tap * p.kb + cb) * b_tile_bytes, &tm_w, bar(kBarW), cb * kBlockK, 0, tap);
        }
        __syncwarp();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                   // previous grid complete: its outputs may be read, ours written
    YMS_PROF_ONLY(const long long prof_t_start = clock64();)
    const uint32_t out_bytes = (uint32_t)(8 * p.th) * 128u;
    const int acc_stride = 512 / p.acc_stages;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
                const Item it = decode_item(p, t);
                const int n0 = it.n_tile * p.block_n;
                for (int cb = 0; cb < p.kb; ++cb) {
                    mbar_wait_acc(bar(kBarAEmpty + as), aph ^ 1u, pw0);
                    mbar_expect_tx(bar(kBarAFull + as), (uint32_t)p.planes * p.halo_bytes);
                    if (p.s2pair) {                                  // the two row-parity planes: 9 pairs x (th+1) rows each
                        for (int py = 0; py < 2; ++py)
                            tma_load_5d(smem_a + (as * 2 + py) * p.halo_stage, &tm_x, bar(kBarAFull + as),
                                        0, it.sx * 8 - 1, py, it.ty * p.th - 1, it.img);
                    } else
                    for (int s = 0; s < p.sub; ++s)
                        tma_load_4d(smem_a + (as * p.sub + s) * p.halo_stage, &tm_x, bar(kBarAFull + as),
                                    cb * kBlockK, (it.sx * p.sub + s) * 8 - 1, it.ty * p.th - 1, it.img);
                    if (++as == p.a_stages) { as = 0; aph ^= 1u; }
                    if (!p.resident) {
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait_acc(bar(kBarBEmpty + bs), bph ^ 1u, pw1);
                            mbar_expect_tx(bar(kBarBFull + bs), (uint32_t)(p.block_n * 128));
                            tma_load_3d(smem_b + bs * b_tile_bytes, &tm_w, bar(kBarBFull + bs), cb * kBlockK, n0, tap);
                            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
                        }
                    }
                }
            }
            YMS_PROF_ONLY(if (prof) { prof[4] = clock64() - prof_t_start; prof[5] = pw0; prof[6] = pw1; })
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // ONE elected thread runs the whole loop: for small N an MMA is only max(N/2, 32 + N/4) cycles of tensor-pipe
        // work (the A operand is re-read from shared memory at 128 B/cycle), so the issue path itself -- barrier wait,
        // fence, descriptor moves, commit -- is what has to be short (scripts/ubench/mma_ring.cu).  Taps and k-steps are
        // fully unrolled (operands = base + immediate); the try_wait of the next ring slot is issued before the MMAs of
        // the current one.
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t hi_a = (1ull << 16) | ((uint64_t)((kHaloPitch * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t hi_b = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t halo16 = (uint32_t)p.halo_stage >> 4;
        const uint32_t btile16 = (uint32_t)b_tile_bytes >> 4;
        if (elect_one()) {
            if (p.resident) { mbar_wait_acc(bar(kBarW), 0u, pw2); tc_fence_after(); }
            YMS_PROF_ONLY(int ntile = 0;)
            const uint32_t a0_16 = (smem_a & 0x3FFFFu) >> 4, b0_16 = (smem_b & 0x3FFFFu) >> 4;
            const uint32_t astage16 = (uint32_t)p.planes * halo16;
            const int tail = ((p.c_in - (p.kb - 1) * kBlockK) + 15) >> 4;
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t a_ready = 0, b_ready = 0;
            for (int t = blockIdx.x; t < p.total_items; t += gridDim.x) {
                mbar_wait_acc(bar(kBarTEmpty + acc), acc_phase ^ 1u, pw3);
                YMS_PROF_ONLY(++ntile;)
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
                for (int cb = 0; cb < p.kb; ++cb) {
                    if (!a_ready) mbar_wait_acc(bar(kBarAFull + as), aph, pw0);
                    tc_fence_after();
                    int nas = as + 1; uint32_t naph = aph;
                    if (nas == p.a_stages) { nas = 0; naph ^= 1u; }
                    a_ready = mbar_try_wait(bar(kBarAFull + nas), naph);
                    const int ksteps = (cb == p.kb - 1) ? tail : 4;
                    const uint32_t a16 = a0_16 + (uint32_t)as * astage16;
                    const uint32_t first = (cb != 0) ? 1u : 0u;
                    const bool last_cb = (cb == p.kb - 1);
                    if (p.s2pair) {
                        const uint64_t hi_a9 = (1ull << 16) | ((uint64_t)((9 * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
                        #pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint32_t row16 = a16 + (ky == 1 ? 0u : halo16) + (uint32_t)((ky == 0 ? 0 : 1) * 9) * 8u;
                            const uint32_t bw = b0_16 + (uint32_t)(2 * ky) * btile16;
                            #pragma unroll
                            for (int k = 0; k < 4; ++k)               // pair x: taps (ky,1) | (ky,2), K = 64
                                umma_bf16(d_tmem, hi_a9 | (uint64_t)(row16 + 8u + 2 * k), hi_b | (uint64_t)(bw + 2 * k), idesc, (ky | k) ? 1u : 0u);
                            #pragma unroll
                            for (int k = 2; k < 4; ++k)               // pair x-1, upper 32 channels: tap (ky,0)
                                umma_bf16(d_tmem, hi_a9 | (uint64_t)(row16 + 2 * k), hi_b | (uint64_t)(bw + btile16 + 2 * k), idesc, 1u);
                        }
                        umma_commit(bar(kBarAEmpty + as));
                        umma_commit(bar(kBarTFull + acc));
                    } else if (p.resident) {
                        const uint32_t b16 = b0_16 + (uint32_t)cb * btile16;
                        const uint32_t bstride16 = (uint32_t)p.kb * btile16;
                        #pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t toff = (uint32_t)((tap / 3) * kHaloPitch + (tap % 3)) * 8u;
                            #pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps) {
                                    // independent accumulators (sub-tiles) are interleaved
                                    #pragma unroll
                                    for (int sidx = 0; sidx < kSub; ++sidx)
                                        umma_bf16(d_tmem + (uint32_t)(sidx * p.block_n), hi_a | (uint64_t)(a16 + sidx * halo16 + toff + 2 * k),
                                                  hi_b | (uint64_t)(b16 + tap * bstride16 + 2 * k), idesc, (tap | k) ? 1u : first);
                                }
                            }
                        }
                        umma_commit(bar(kBarAEmpty + as));
                        if (last_cb) umma_commit(bar(kBarTFull + acc));
                    } else {
                        #pragma unroll 1
                        for (int tap = 0; tap < 9; ++tap) {
                            if (!b_ready) mbar_wait_acc(bar(kBarBFull + bs), bph, pw1);
                            tc_fence_after();
                            int nbs = bs + 1; uint32_t nbph = bph;
                            if (nbs == p.b_stages) { nbs = 0; nbph ^= 1u; }
                            b_ready = mbar_try_wait(bar(kBarBFull + nbs), nbph);
                            const uint32_t b16 = b0_16 + (uint32_t)bs * btile16;
                            const uint32_t toff = (uint32_t)((tap / 3) * kHaloPitch + (tap % 3)) * 8u;
                            const uint32_t accf = tap ? 1u : first;
                            #pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps) {
                                    #pragma unroll
                                    for (int sidx = 0; sidx < (kSub >= 2 ? 2 : 1); ++sidx)
                                        umma_bf16(d_tmem + (uint32_t)(sidx * p.block_n), hi_a | (uint64_t)(a16 + sidx * halo16 + toff + 2 * k),
                                                  hi_b | (uint64_t)(b16 + 2 * k), idesc, k ? 1u : accf);
                                }
                            }
                            umma_commit(bar(kBarBEmpty + bs));
                            if (tap == 8) {
                                umma_commit(bar(kBarAEmpty + as));
                                if (last_cb) umma_commit(bar(kBarTFull + acc));
                            }
                            bs = nbs; bph = nbph;
                        }
                    }
                    as = nas; aph = naph;
                }
                if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
            YMS_PROF_ONLY(if (prof) { prof[0] = clock64() - prof_t_start; prof[1] = pw0; prof[2] = pw1; prof[3] = pw3; prof[10] = pw2;
                                      prof[11] = ntile; prof[9] = prof_t_start - prof_t_entry; })
        }
        __syncwarp();
    } else {
        // ================= epilogue: up to 4 groups of 4 warps, group e drains accumulator stage e =================
        const int grp = (warp - 2) >> 2;
        const int gps = kEpiGroups / p.acc_stages;         // groups sharing one accumulator stage
        const int stage_id = grp / gps, sub_id = grp - stage_id * gps;
        bias_stage(s_bias, bias_regs, p.bias, p.c_out, p.bias_pad, p.act);
        {
            EpiShared e;
            e.tm_y = &tm_y; e.tm_res = &tm_res;
            e.res_bar = bar(kBarRes + grp);
            e.s_out = smem_out0 + grp * kStageOutBytes;
            e.s_bias = s_bias;
            e.block_n = p.block_n; e.c_out = p.c_out; e.act = p.act; e.has_res = p.has_res;
            e.out_bytes = out_bytes;
            e.bar_id = 1 + grp;
            e.leader = ((warp - 2) & 3) == 0 && lane == 0;
            e.row = (warp & 3) * 32 + lane;
            const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(stage_id * acc_stride);
            const int n_chunks = (p.block_n + 63) >> 6;
            uint32_t res_phase = 0u, acc_phase = 0u;
            for (int t = blockIdx.x + stage_id * gridDim.x; t < p.total_items; t += p.acc_stages * gridDim.x) {
                const Item it = decode_item(p, t);
                mbar_wait_acc(bar(kBarTFull + stage_id), acc_phase, pw0);
                acc_phase ^= 1u;
                tc_fence_after();
                int unit = 0;                                        // (sub-tile, chunk) units dealt round-robin to the groups
                for (int sidx = 0; sidx < p.sub; ++sidx) {
                    EpiTile tl;
                    tl.n0 = it.n_tile * p.block_n; tl.x0 = (it.sx * p.sub + sidx) * 8; tl.y0 = it.ty * p.th; tl.img = it.img;
                    if (tl.x0 >= p.out_w) continue;                  // sub-tile entirely outside the image
                    for (int ch = 0; ch < n_chunks; ++ch, ++unit)
                        if (unit % gps == sub_id) epilogue_chunk_bf16(e, res_phase, t_lane + (uint32_t)(sidx * p.block_n), tl, ch);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(kBarTEmpty + stage_id));
            }
            if (e.leader) tma_store_wait_read<0>();
            YMS_PROF_ONLY(if (prof && warp == 2 && lane == 0) { prof[7] = clock64() - prof_t_start; prof[8] = pw0; })
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    YMS_PROF_ONLY(if (prof && threadIdx.x == 0) prof[12] = clock64() - prof_t_entry;)
}


// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (yms_conv_params.variant == 5).  At N = 128 one tcgen05.mma (M = 128, K = 16) reads 4 KB of A and 4 KB
// of B from shared memory per 64 cycles of tensor-pipe work -- exactly the 128 B/cycle the SM has, so every other byte that
// moves through shared memory (TMA writes, the epilogue's staging) stretches the MMAs (role accounting: ~134 cycles per
// MMA in-kernel against 64 in isolation).  Two CTAs of a cluster on the two SMs of a TPC instead run ONE MMA of M = 256:
// each supplies the halo tile of its own 8 x th sub-tile (the cluster owns two x-adjacent sub-tiles) and HALF of the weight
// tile (N/2 rows), the tensor cores exchange the halves: 6 KB per 64 cycles per SM, and half the weight traffic from L2.
// Accumulator rows 0..127 / 128..255 live in the TMEM of CTA 0 / CTA 1: the epilogue is the single-CTA one, per CTA.
// Protocol: the leader (rank 0) issues all MMAs; the full barriers are the leader's and are signalled by the TMA loads of both
// CTAs (cta_group::2 loads, expect_tx of the pair's bytes by the leader's producer); tcgen05.commit arrives on the empty /
// accumulator-full barriers of both CTAs (multicast); the epilogue groups of both CTAs arrive on the leader's
// accumulator-empty barrier.  Same per-CTA shared-memory layout and ring discipline as conv3x3_kernel<1>.
//
// kVy, virtual-row tiling (variant 7) for maps whose height is not a multiple of 16 (40 -> 3 bands of 14 rows, 20 -> 2 of 10: up to
// 37 % of every MMA's rows are padding, plus a sixth sub-tile column for five when pairing in x).  All images are stacked into ONE
// column of virtual rows, vh = H + 2 apart -- the two extra rows are the zero padding below one image and above the next -- and
// that column is cut into bands of exactly 16 rows; a cluster takes one 8-pixel column of two consecutive bands.  A band inside
// one image loads its halo with the usual single box (rows H, H + 1 are out of bounds = zero-filled).  A band that crosses into the
// next image loads its 18 halo rows two by two through a two-row box (tm_xr; H is even, so a row pair never spans two images): two boxes would zero-fill each other.  The two dummy
// output rows are computed and dropped: the epilogue stores the tile as a box into the first image (TMA clips the rows past its end) and
// the next image's rows two by two (a box at a negative row is rejected by the hardware).
template <bool kVy>
__global__ void __launch_bounds__(kThreads3, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_res,
                    const __grid_constant__ CUtensorMap tm_xr, const __grid_constant__ CUtensorMap tm_rr,
                    const __grid_constant__ CUtensorMap tm_yr, const __grid_constant__ Conv3Params p) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    long long pw0 = 0, pw1 = 0, pw2 = 0, pw3 = 0;   // wait-cycle accumulators (dead code unless -DYMS_PROF)
    (void)pw0; (void)pw1; (void)pw2; (void)pw3;
    YMS_PROF_ONLY(const long long prof_t_entry = clock64(); long long* prof = (p.prof && blockIdx.x < kNumSMs) ? p.prof + 16 * blockIdx.x : nullptr;)
    const uint32_t base = smem_u32(smem_dyn);
    unsigned char* gbase = smem_dyn;
    if (base & 1023u) __trap();
    const uint32_t rank = cluster_ctarank();
    const bool leader_cta = rank == 0;
    const int half_n = p.block_n >> 1;
    const int b_tile_bytes = (half_n * 128 + 1023) & ~1023;          // this CTA's half of a weight tile
    const int a_region = ((p.a_stages * p.halo_stage) + 1023) & ~1023;
    const int b_region = (p.resident ? 9 * p.kb : 3 * p.b_stages) * b_tile_bytes;     // streamed: one ring slot = the three taps of a kernel row
    const uint32_t smem_a = base;
    const uint32_t smem_b = base + a_region;
    const uint32_t smem_out0 = smem_b + b_region;
    unsigned char* g_out0 = gbase + a_region + b_region;
    float* s_bias = reinterpret_cast<float*>(g_out0 + p.epi_groups * kStageOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int slot) { return bar0 + 8u * slot; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int gps = p.epi_groups / p.acc_stages;           // epilogue groups sharing one accumulator stage

    const BiasRegs bias_regs = bias_fetch(p.bias, p.c_out);          // staged by the epilogue warps after the CTA-wide sync
    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tm_x); prefetch_tmap(&tm_w); prefetch_tmap(&tm_y);
            if (p.has_res) prefetch_tmap(&tm_res);
            if (kVy) { prefetch_tmap(&tm_xr); prefetch_tmap(&tm_yr); if (p.has_res) prefetch_tmap(&tm_rr); }
        }
        for (int i = lane; i < kNumBars; i += 32) {
            const bool is_tempty = i >= kBarTEmpty && i < kBarTEmpty + 4;
            mbar_init(bar(i), is_tempty ? 2 * 4 * gps : 1);          // accumulator-empty (leader's): the epilogue warps of BOTH CTAs
        }
        fence_barrier_init();
        __syncwarp();
    }
    cluster_sync_all();                                    // the peer's barriers exist before anything signals them
    if (warp == 0 && p.resident && elect_one()) {
        // each CTA fetches ITS half of every weight tile; the leader's barrier collects the bytes of both
        if (leader_cta) mbar_expect_tx(bar(kBarW), 2u * (uint32_t)(9 * p.kb) * (uint32_t)(half_n * 128));
        for (int tap = 0; tap < 9; ++tap)
            for (int cb = 0; cb < p.kb; ++cb)
                tma_load_3d_2sm(smem_b + (tap * p.kb + cb) * b_tile_bytes, &tm_w, bar(kBarW), cb * kBlockK, (int)rank * half_n, tap);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t out_bytes = (uint32_t)(8 * p.th) * 128u;
    const int acc_stride = 512 / p.acc_stages;
    const int cid = (int)cluster_id_x(), ncl = (int)num_clusters_x();
    YMS_PROF_ONLY(const long long prof_t_start = clock64();)
    // kVy: this CTA's sub-tile of item t = column sx, band 2 * (t / super_x) + rank; the band's first virtual row = row y0 of image img
    struct VTile { int sx, img, y0; };
    auto vtile = [&](int t) {
        VTile v;
        const uint32_t bp = fast_div((uint32_t)t, p.mg_super_x);
        v.sx = t - (int)bp * p.super_x;
        const uint32_t v0 = (2u * bp + rank) * 16u;
        v.img = (int)fast_div(v0, p.mg_vh);
        v.y0 = (int)v0 - v.img * p.vh;
        return v;
    };

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own halo tile, own half of the weights) =================
        if (elect_one()) {
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            for (int t = cid; t < p.total_items; t += ncl) {
                const Item it = kVy ? Item{} : decode_item(p, t);
                const VTile vt = kVy ? vtile(t) : VTile{};
                for (int cb = 0; cb < p.kb; ++cb) {
                    mbar_wait_acc(bar(kBarAEmpty + as), aph ^ 1u, pw0);
                    if (leader_cta) mbar_expect_tx(bar(kBarAFull + as), 2u * p.halo_bytes);
                    if (!kVy) {
                        tma_load_4d_2sm(smem_a + as * p.halo_stage, &tm_x, bar(kBarAFull + as),
                                        cb * kBlockK, (it.sx * 2 + (int)rank) * 8 - 1, it.ty * p.th - 1, it.img);
                    } else if (vt.y0 + 16 <= p.vh - 1) {             // halo rows y0 - 1 .. y0 + 16 <= H + 1: one image, rows >= H read as zeros
                        tma_load_4d_2sm(smem_a + as * p.halo_stage, &tm_x, bar(kBarAFull + as), cb * kBlockK, vt.sx * 8 - 1, vt.y0 - 1, vt.img);
                    } else {
                        // TWO rows per box: H is even, so the bands (16 rows) and the images (vh rows) start at even virtual rows and
                        // a pair of halo rows (odd, even) never spans two images; rows H, H + 1 and images >= batch are zero fill
                        for (int r = 0; r < 18; r += 2) {
                            const int y = vt.y0 - 1 + r, wrap = y >= p.vh - 1 ? 1 : 0;       // virtual row vh - 1 is row -1 of the next image
                            tma_load_4d_2sm(smem_a + as * p.halo_stage + r * (kHaloPitch * 128), &tm_xr, bar(kBarAFull + as),
                                            cb * kBlockK, vt.sx * 8 - 1, y - wrap * p.vh, vt.img + wrap);
                        }
                    }
                    if (++as == p.a_stages) { as = 0; aph ^= 1u; }
                    if (!p.resident) {
                        for (int ky = 0; ky < 3; ++ky) {             // a slot = the 3 taps of a kernel row: a third of the ring hand-shakes
                            mbar_wait_acc(bar(kBarBEmpty + bs), bph ^ 1u, pw1);
                            if (leader_cta) mbar_expect_tx(bar(kBarBFull + bs), 6u * (uint32_t)(half_n * 128));
                            for (int kx = 0; kx < 3; ++kx)
                                tma_load_3d_2sm(smem_b + (bs * 3 + kx) * b_tile_bytes, &tm_w, bar(kBarBFull + bs), cb * kBlockK, (int)rank * half_n, ky * 3 + kx);
                            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
                        }
                    }
                }
            }
            YMS_PROF_ONLY(if (prof) { prof[4] = clock64() - prof_t_start; prof[5] = pw0; prof[6] = pw1; })
        }
    } else if (warp == 1) {
        // ================= MMA issuer: one elected thread of the LEADER =================
        if (leader_cta && elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint64_t hi_a = (1ull << 16) | ((uint64_t)((kHaloPitch * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint64_t hi_b = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t halo16 = (uint32_t)p.halo_stage >> 4;
            const uint32_t btile16 = (uint32_t)b_tile_bytes >> 4;
            if (p.resident) { mbar_wait_acc(bar(kBarW), 0u, pw2); tc_fence_after(); }
            const uint32_t a0_16 = (smem_a & 0x3FFFFu) >> 4, b0_16 = (smem_b & 0x3FFFFu) >> 4;
            const int tail = ((p.c_in - (p.kb - 1) * kBlockK) + 15) >> 4;
            int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t a_ready = 0, b_ready = 0;
            for (int t = cid; t < p.total_items; t += ncl) {
                mbar_wait_acc(bar(kBarTEmpty + acc), acc_phase ^ 1u, pw3);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
                for (int cb = 0; cb < p.kb; ++cb) {
                    if (!a_ready) mbar_wait_acc(bar(kBarAFull + as), aph, pw0);
                    tc_fence_after();
                    int nas = as + 1; uint32_t naph = aph;
                    if (nas == p.a_stages) { nas = 0; naph ^= 1u; }
                    a_ready = mbar_try_wait(bar(kBarAFull + nas), naph);
                    const int ksteps = (cb == p.kb - 1) ? tail : 4;
                    const uint32_t a16 = a0_16 + (uint32_t)as * halo16;
                    const uint32_t first = (cb != 0) ? 1u : 0u;
                    const bool last_cb = (cb == p.kb - 1);
                    if (p.resident) {
                        const uint32_t b16 = b0_16 + (uint32_t)cb * btile16;
                        const uint32_t bstride16 = (uint32_t)p.kb * btile16;
                        #pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            const uint32_t toff = (uint32_t)((tap / 3) * kHaloPitch + (tap % 3)) * 8u;
                            #pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < ksteps)
                                    umma_bf16_2sm(d_tmem, hi_a | (uint64_t)(a16 + toff + 2 * k), hi_b | (uint64_t)(b16 + tap * bstride16 + 2 * k),
                                                  idesc, (tap | k) ? 1u : first);
                        }
                        umma_commit_2sm(bar(kBarAEmpty + as));
                        if (last_cb) umma_commit_2sm(bar(kBarTFull + acc));
                    } else {
                        #pragma unroll 1
                        for (int ky = 0; ky < 3; ++ky) {
                            if (!b_ready) mbar_wait_acc(bar(kBarBFull + bs), bph, pw1);
                            tc_fence_after();
                            int nbs = bs + 1; uint32_t nbph = bph;
                            if (nbs == p.b_stages) { nbs = 0; nbph ^= 1u; }
                            b_ready = mbar_try_wait(bar(kBarBFull + nbs), nbph);
                            #pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint32_t b16 = b0_16 + (uint32_t)(bs * 3 + kx) * btile16;
                                const uint32_t toff = (uint32_t)(ky * kHaloPitch + kx) * 8u;
                                #pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (k < ksteps)
                                        umma_bf16_2sm(d_tmem, hi_a | (uint64_t)(a16 + toff + 2 * k), hi_b | (uint64_t)(b16 + 2 * k), idesc,
                                                      (ky | kx | k) ? 1u : first);
                            }
                            umma_commit_2sm(bar(kBarBEmpty + bs));
                            if (ky == 2) {
                                umma_commit_2sm(bar(kBarAEmpty + as));
                                if (last_cb) umma_commit_2sm(bar(kBarTFull + acc));
                            }
                            bs = nbs; bph = nbph;
                        }
                    }
                    as = nas; aph = naph;
                }
                if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
            YMS_PROF_ONLY(if (prof) { prof[0] = clock64() - prof_t_start; prof[1] = pw0; prof[2] = pw1; prof[3] = pw3; prof[10] = pw2;
                                      prof[11] = (p.total_items - cid + ncl - 1) / ncl; prof[9] = prof_t_start - prof_t_entry; })
        }
        __syncwarp();
    } else {
        // ================= epilogue (both CTAs, each its own 128 accumulator rows = its own sub-tile) =================
        const int grp = (warp - 2) >> 2;
        const int stage_id = grp / gps, sub_id = grp - stage_id * gps;
        bias_stage(s_bias, bias_regs, p.bias, p.c_out, p.bias_pad, p.act);
        EpiShared e;
        e.tm_y = &tm_y; e.tm_res = &tm_res;
        e.res_bar = bar(kBarRes + grp);
        e.s_out = smem_out0 + grp * kStageOutBytes;
        e.s_bias = s_bias;
        e.block_n = p.block_n; e.c_out = p.c_out; e.act = p.act; e.has_res = p.has_res;
        e.out_bytes = out_bytes;
        e.bar_id = 1 + grp;
        e.leader = ((warp - 2) & 3) == 0 && lane == 0;
        e.row = (warp & 3) * 32 + lane;
        e.tm_res_row = &tm_rr; e.tm_y_row = &tm_yr;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(stage_id * acc_stride);
        const int n_chunks = (p.block_n + 63) >> 6;
        uint32_t res_phase = 0u, acc_phase = 0u;
        for (int t = cid + stage_id * ncl; t < p.total_items; t += p.acc_stages * ncl) {
            EpiTile tl;
            if (kVy) {
                const VTile vt = vtile(t);
                tl.n0 = 0; tl.x0 = vt.sx * 8; tl.y0 = vt.y0; tl.img = vt.img; tl.vh = p.vh;
            } else {
                const Item it = decode_item(p, t);
                tl.n0 = 0; tl.x0 = (it.sx * 2 + (int)rank) * 8; tl.y0 = it.ty * p.th; tl.img = it.img;
            }
            mbar_wait_acc(bar(kBarTFull + stage_id), acc_phase, pw0);
            acc_phase ^= 1u;
            tc_fence_after();
            if (kVy ? (tl.img < p.batch && (tl.y0 < p.out_h || tl.img + 1 < p.batch))      // (a band past the last image / of dummy rows only)
                    : (tl.x0 < p.out_w))                             // (the pair's second sub-tile may lie outside the image)
                for (int ch = sub_id; ch < n_chunks; ch += gps) epilogue_chunk_bf16<false, kVy ? 1 : 0>(e, res_phase, t_lane, tl, ch);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar(kBarTEmpty + stage_id), 0u);
        }
        if (e.leader) tma_store_wait_read<0>();
        YMS_PROF_ONLY(if (prof && warp == 2 && lane == 0) { prof[7] = clock64() - prof_t_start; prof[8] = pw0; })
    }

    tc_fence_before();
    __syncthreads();
    YMS_PROF_ONLY(if (prof && threadIdx.x == 0) prof[12] = clock64() - prof_t_entry;)
    cluster_sync_all();                                    // nobody leaves (or frees TMEM) while the peer may still signal or be read
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
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
    // variant 7: the pair kernel with virtual-row tiling (bands of exactly 16 rows over the stacked images, see conv3x3_pair_kernel)
    k.vy = (q->variant == 7) ? 1 : 0;
    if (k.vy) {
        k.th = 16; k.vh = H + 2;
        k.bands = ceil_div(q->batch * k.vh, 16);
        k.mg_vh = fast_div_magic(k.vh);
    }
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
    k.desc_mode = 0;
    // variant 5: CTA-pair kernel (conv3x3_pair_kernel): two x-adjacent sub-tiles per cluster, half a weight tile per CTA
    k.pair = (q->variant == 5 || k.vy) ? 1 : 0;
    if (k.pair && (k.n_tiles != 1 || (!k.vy && k.tiles_x < 2))) return fail(YMS_E_UNSUPPORTED, "conv3x3 (variant 5): needs c_out <= 256 and a map at least 9 pixels wide");
    if (k.vy && (k.bands < 2 || H < 16 || (H & 1)))
        return fail(YMS_E_UNSUPPORTED, "conv3x3 (variant 7): needs maps of even height, at least 16 rows (a band may span two images, not three; two-row boxes)");

    const int b_tile = ((k.pair ? k.block_n / 2 : k.block_n) * 128 + 1023) & ~1023;
    const int halo_stage = (int)k.halo_bytes;
    k.halo_stage = halo_stage;
    int fixed = kEpiGroups * kStageOutBytes + k.bias_pad * 4 + kNumBars * 8 + 16;
    const int resident_bytes = 9 * k.kb * b_tile;
    // largest ring depth (in items of `sub` halos) that fits next to `other` bytes of weights
    auto max_a_stages = [&](int sub, int other) {
        int n = 0;
        while (n < kRing && ((((n + 1) * sub * halo_stage + 1023) & ~1023) + other + fixed <= kSmemLimit3)) ++n;
        return n;
    };
    k.epi_groups = kEpiGroups;
    k.resident = (k.n_tiles == 1 && max_a_stages(1, resident_bytes) >= 2) ? 1 : 0;
    if (k.pair && !k.resident && k.n_tiles == 1) {
        // half tiles per CTA: with TWO epilogue groups (32 KB of staging instead of 64) the weights of a 128 -> 128 layer
        // (9 x 2 x 8 KB) stay resident next to a two-deep halo ring -- no weight traffic from L2, no weight ring hand-shakes
        fixed -= 2 * kStageOutBytes;
        if (max_a_stages(1, resident_bytes) >= 2) { k.resident = 1; k.epi_groups = 2; }
        else fixed += 2 * kStageOutBytes;
    }
    if (k.pair) {
        k.sub = 1;                                         // per CTA; the cluster covers two sub-tiles
        k.b_stages = k.resident ? 0 : 3;                   // ring slots of THREE half tiles (the taps of one kernel row)
        for (;;) {
            k.a_stages = max_a_stages(1, k.resident ? resident_bytes : 3 * k.b_stages * b_tile);
            if (k.a_stages >= 2 || k.resident || k.b_stages == 2) break;
            --k.b_stages;
        }
    } else if (k.resident) {
        // two sub-tiles per item when they fit: their MMAs use independent accumulators and are
        // interleaved, which hides part of the ~100-cycle per-instruction tcgen05.mma latency
        const long long sub_tiles = (long long)k.tiles_x * k.tiles_y * k.batch;
        k.sub = (k.tiles_x >= 2 && sub_tiles >= 4096) ? 2 : 1;                       // pairing only pays with many tiles per CTA
        if (q->variant == 2) k.sub = 1; else if (q->variant == 3 && k.tiles_x >= 2) k.sub = 2;
        k.b_stages = 0;
        if ((k.sub != 1 && k.sub != 2 && k.sub != 4) || k.sub * k.block_n > 512) k.sub = 1;
        k.a_stages = max_a_stages(k.sub, resident_bytes);
        if (k.a_stages < 2) { k.sub = 1; k.a_stages = max_a_stages(1, resident_bytes); }
    } else {
        k.sub = (k.tiles_x >= 2 && 2 * k.block_n <= 512) ? 2 : 1;
        if (q->variant == 2) k.sub = 1;
        k.b_stages = 4;
        for (;;) {
            k.a_stages = max_a_stages(k.sub, k.b_stages * b_tile);
            if (k.a_stages >= 2 || k.b_stages == 2) break;
            --k.b_stages;
        }
        if (k.a_stages < 2 && k.sub == 2) {
            k.sub = 1; k.b_stages = 4;
            k.a_stages = max_a_stages(1, k.b_stages * b_tile);
        }
    }
    if (k.a_stages > kRing) k.a_stages = kRing;
    if (k.a_stages < 2) return fail(YMS_E_UNSUPPORTED, "conv3x3: tile does not fit in shared memory");
    k.acc_stages = (k.sub * k.block_n <= 128) ? 4 : ((k.sub * k.block_n <= 256) ? 2 : 1);
    if (k.acc_stages > k.epi_groups) k.acc_stages = k.epi_groups;
    k.planes = k.sub; k.wtiles = 9; k.pitch = kHaloPitch; k.s2pair = 0;
    k.super_x = ceil_div(k.tiles_x, k.pair ? 2 : k.sub);
    k.total_items = k.super_x * k.tiles_y * k.batch * k.n_tiles;
    if (k.vy) { k.super_x = k.tiles_x; k.total_items = k.tiles_x * ceil_div(k.bands, 2); }     // item = one sub-tile column x two consecutive bands
    k.mg_n_tiles = fast_div_magic(k.n_tiles); k.mg_super_x = fast_div_magic(k.super_x); k.mg_tiles_y = fast_div_magic(k.tiles_y);
    pl->grid = k.total_items < kNumSMs ? k.total_items : kNumSMs;
    if (k.pair) pl->grid = 2 * (k.total_items < kNumSMs / 2 ? k.total_items : kNumSMs / 2);          // clusters of two CTAs
    pl->smem = (size_t)((k.a_stages * k.sub * halo_stage + 1023) & ~1023) + (size_t)(k.resident ? resident_bytes : (k.pair ? 3 : 1) * k.b_stages * b_tile) + fixed;

    int rc;
    if ((rc = encode_act(&pl->tm_x, q->x, q->c_in, q->x_pixel_stride, q->batch, H, W, false, kHaloPitch, k.th + 2, 1, "x(halo)"))) return rc;
    {
        uint64_t dims[3] = {(uint64_t)q->c_in, (uint64_t)q->c_out, 9};
        uint64_t strides[2] = {(uint64_t)q->c_in * 2, (uint64_t)q->c_in * 2 * (uint64_t)q->c_out};
        uint32_t box[3] = {kBlockK, (uint32_t)(k.pair ? k.block_n / 2 : k.block_n), 1};
        uint32_t es[3] = {1, 1, 1};
        if ((rc = encode_map(&pl->tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->weight, dims, strides, box, es, "w"))) return rc;
    }
    if ((rc = encode_act(&pl->tm_y, q->y, q->c_out, q->y_pixel_stride, q->batch, H, W, false, 8, k.th, 1, "y"))) return rc;
    if (q->residual) {
        if ((rc = encode_act(&pl->tm_res, q->residual, q->c_out, q->res_pixel_stride, q->batch, H, W, false, 8, k.th, 1, "res"))) return rc;
    } else pl->tm_res = pl->tm_y;
    pl->tm_x2 = pl->tm_x;
    pl->tm_res2 = pl->tm_res; pl->tm_y2 = pl->tm_y;
    if (k.vy) {
        // two-row boxes for the bands that cross an image boundary
        if ((rc = encode_act(&pl->tm_y2, q->y, q->c_out, q->y_pixel_stride, q->batch, H, W, false, 8, 2, 1, "y(rows)"))) return rc;
        if ((rc = encode_act(&pl->tm_x2, q->x, q->c_in, q->x_pixel_stride, q->batch, H, W, false, kHaloPitch, 2, 1, "x(halo rows)"))) return rc;
        if (q->residual && (rc = encode_act(&pl->tm_res2, q->residual, q->c_out, q->res_pixel_stride, q->batch, H, W, false, 8, 2, 1, "res(rows)"))) return rc;
    }

    const double m = (double)q->batch * H * W;
    pl->flops = 2.0 * m * q->c_out * (double)q->c_in * 9.0;
    pl->bytes = 2.0 * m * q->c_in + 2.0 * m * q->c_out + 2.0 * 9.0 * q->c_out * q->c_in + (q->residual ? 2.0 * m * q->c_out : 0.0);

    static std::atomic<unsigned long long> attr_seen{0};
    if (first_use_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv3x3_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
        if (e != cudaSuccess) return fail((int)e, "conv3x3: smem attribute: %s", cudaGetErrorString(e));
    }
    return 0;
}

// Stride-2 pair-line mode (see the header comment).  q->weight is PAIR-PACKED: bf16 [6][c_out][64], tile 2*ky =
// [w(ky,1) | w(ky,2)], tile 2*ky+1 = [0 | w(ky,0)] (32 input channels each).
int conv3_s2pair_plan_init(yms_conv_plan* pl, const yms_conv_params* q) {
    if (!(q->ksize == 3 && q->stride == 2 && q->c_in == 32 && q->c_in2 == 0 && q->x_pixel_stride == 32 && q->out_dtype == YMS_DTYPE_BF16 &&
          q->c_out <= 256 && !q->residual))
        return fail(YMS_E_UNSUPPORTED, "conv (variant 4): needs 3x3/s2, c_in == 32, dense input, bf16 output, c_out <= 256, no residual");
    Conv3Params& k = pl->k3;
    memset(&k, 0, sizeof(k));
    pl->kind = 1;
    const int H = q->in_h / 2, W = q->in_w / 2;                 // output size
    k.out_w = W; k.out_h = H; k.batch = q->batch;
    k.c_in = 64; k.c_out = q->c_out; k.kb = 1;
    k.n_tiles = 1; k.block_n = ((q->c_out + 15) / 16) * 16;
    k.act = q->act ? 1 : 0; k.has_res = 0;
    k.bias_pad = k.block_n + 64;
    k.bias = q->bias;
    k.s2pair = 1; k.planes = 2; k.wtiles = 6; k.pitch = 9; k.sub = 1; k.resident = 1; k.b_stages = 0;
    const int b_tile = (k.block_n * 128 + 1023) & ~1023;
    const int fixed = kEpiGroups * kStageOutBytes + k.bias_pad * 4 + kNumBars * 8 + 16;
    k.th = 0;
    for (int th_max = 16; th_max >= 4 && !k.th; th_max -= 2) {   // tallest tile that still leaves a 3-deep ring
        const int ny = ceil_div(H, th_max), th = ceil_div(H, ny);
        const int stage = 2 * 9 * (th + 1) * 128;
        if (((3 * stage + 1023) & ~1023) + 6 * b_tile + fixed <= kSmemLimit3) k.th = th;
    }
    if (!k.th) return fail(YMS_E_UNSUPPORTED, "conv (variant 4): does not fit in shared memory");
    k.halo_bytes = (uint32_t)(9 * (k.th + 1) * 128);
    k.halo_stage = (int)k.halo_bytes;
    k.a_stages = 0;
    while (k.a_stages < kRing && ((((k.a_stages + 1) * 2 * k.halo_stage + 1023) & ~1023) + 6 * b_tile + fixed <= kSmemLimit3)) ++k.a_stages;
    k.tiles_x = ceil_div(W, 8); k.tiles_y = ceil_div(H, k.th);
    k.acc_stages = (k.block_n <= 128) ? 4 : 2;
    k.super_x = k.tiles_x;
    k.total_items = k.super_x * k.tiles_y * k.batch;
    k.mg_n_tiles = fast_div_magic(1); k.mg_super_x = fast_div_magic(k.super_x); k.mg_tiles_y = fast_div_magic(k.tiles_y);
    pl->grid = k.total_items < kNumSMs ? k.total_items : kNumSMs;
    pl->smem = (size_t)((k.a_stages * 2 * k.halo_stage + 1023) & ~1023) + (size_t)6 * b_tile + fixed;

    int rc;
    {   // dense stride-2 input as (2C = 64 channels, W/2 pairs, 2 row parities, H/2, N)
        const uint64_t Wi = (uint64_t)q->in_w, Hi = (uint64_t)q->in_h;
        uint64_t dims[5] = {64, Wi / 2, 2, Hi / 2, (uint64_t)q->batch};
        uint64_t strides[4] = {64 * 2, Wi * 32 * 2, 2 * Wi * 32 * 2, Hi * Wi * 32 * 2};
        uint32_t box[5] = {64, 9, 1, (uint32_t)(k.th + 1), 1};
        uint32_t es[5] = {1, 1, 1, 1, 1};
        if ((rc = encode_map(&pl->tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, q->x, dims, strides, box, es, "x(s2 pair planes)"))) return rc;
    }
    {
        uint64_t dims[3] = {64, (uint64_t)q->c_out, 6};
        uint64_t strides[2] = {64 * 2, 64 * 2 * (uint64_t)q->c_out};
        uint32_t box[3] = {kBlockK, (uint32_t)k.block_n, 1};
        uint32_t es[3] = {1, 1, 1};
        if ((rc = encode_map(&pl->tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->weight, dims, strides, box, es, "w(pair-packed)"))) return rc;
    }
    if ((rc = encode_act(&pl->tm_y, q->y, q->c_out, q->y_pixel_stride, q->batch, H, W, false, 8, k.th, 1, "y"))) return rc;
    pl->tm_res = pl->tm_y;
    pl->tm_x2 = pl->tm_x;
    const double m = (double)q->batch * H * W;
    pl->flops = 2.0 * m * q->c_out * 32.0 * 9.0;
    pl->bytes = 2.0 * (double)q->batch * q->in_h * q->in_w * 32.0 + 2.0 * m * q->c_out + 2.0 * 9.0 * q->c_out * 32.0;
    cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit3);
    if (e != cudaSuccess) return fail((int)e, "conv3x3: smem attribute: %s", cudaGetErrorString(e));
    return 0;
}

int conv3_plan_run(const yms_conv_plan* pl0, cudaStream_t stream) {
    yms_conv_plan plc = *pl0;                  // by-value launch parameters; the debug hook patches the counter buffer in
    plc.k3.prof = g_prof_buf;
    const yms_conv_plan* pl = &plc;
    cudaError_t le;
    if (pl->k3.pair) le = launch_pdl_cluster(pl->k3.vy ? conv3x3_pair_kernel<true> : conv3x3_pair_kernel<false>, pl->grid, 64 + pl->k3.epi_groups * kEpiGroupThreads, pl->smem, stream, 2,
                                             pl->tm_x, pl->tm_w, pl->tm_y, pl->tm_res, pl->tm_x2, pl->tm_res2, pl->tm_y2, pl->k3);
    else if (pl->k3.sub == 1) le = launch_pdl(conv3x3_kernel<1>, pl->grid, kThreads3, pl->smem, stream, pl->tm_x, pl->tm_w, pl->tm_y, pl->tm_res, pl->k3);
    else if (pl->k3.sub == 2) le = launch_pdl(conv3x3_kernel<2>, pl->grid, kThreads3, pl->smem, stream, pl->tm_x, pl->tm_w, pl->tm_y, pl->tm_res, pl->k3);
    else le = launch_pdl(conv3x3_kernel<4>, pl->grid, kThreads3, pl->smem, stream, pl->tm_x, pl->tm_w, pl->tm_y, pl->tm_res, pl->k3);
    if (le != cudaSuccess) return fail((int)le, "conv3x3_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("conv3x3_kernel");
}

}  // namespace yms
