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
#include "decode_math.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <new>

namespace yms {
namespace {

constexpr int kBlockM = 128;
constexpr int kMaxStages = 8;
constexpr int kATileBytes = kBlockM * kBlockK * 2;      // 16 KB
using namespace tc;
constexpr int kThreads = kConvThreads;
constexpr int kSmemLimitHalf = 115712;         // 113 KB: two CTAs (+ 1 KB reserved each) fill the SM's 228 KB
constexpr int kSmemLimit = 232448;             // 227 KB

struct TileCoord { int n_tile, img, x0, y0; };
__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int t) {
    TileCoord c;
    uint32_t m = fast_div((uint32_t)t, p.mg_n_tiles);
    c.n_tile = t - (int)m * p.n_tiles;
    uint32_t q = fast_div(m, p.mg_tiles_x);
    const int tx = (int)(m - q * p.tiles_x);
    m = q;
    q = fast_div(m, p.mg_tiles_y);
    const int ty = (int)(m - q * p.tiles_y);
    c.img = (int)q;
    c.x0 = tx * p.tw; c.y0 = ty * p.th;
    return c;
}

// Decode-fused epilogue of a head branch's final (biased, linear) 1x1 convolution.  The tile is 128 consecutive pixels of
// the flat (B*H*W) view; accumulator row `row` = TMEM lane = ONE anchor, its columns = the branch's logits, so the DFL
// softmax-expectation (box branch) and the sigmoid scores + first-max class (class branch) are thread-local: the fp32
// logits never travel to HBM and back (yolov8_head.py:127-144, components.py:176-191, tools/test.py:166-179).  The
// arithmetic is decode_math.cuh, i.e. bit-identical to head_decode_v2_kernel on the logits this conv would have stored.
// Every lane executes the (warp-collective) tcgen05.ld; only the stores are predicated on the row being a real pixel.
//
// Stores: a prediction row is 4 + nc floats, so "one thread stores its row" scatters every warp store over 32 lines (first
// version: the class branch at 80x80 took 32 us instead of 19 -- one 16-byte packet per lane on the SM -> L2 path).  The
// class scores therefore go through the warp's 32 rows of the group's staging tile (128 B = 32 scores per row, swizzled
// like the TMA path) and leave as 128-byte row segments: 8 lanes per row, 4 rows per store instruction.
__device__ __forceinline__ void epilogue_decode(const DecodeFuse& d, const float* s_bias, int nc, uint32_t t_row, uint32_t s_out,
                                                int row, int lane, uint32_t m, uint32_t m_total) {
    int arow = -1, pix = 0;                                       // prediction row of this thread's pixel (-1: past the last pixel)
    if (m < m_total) {
        const uint32_t img = m / (uint32_t)d.hw;
        pix = (int)(m - img * (uint32_t)d.hw);
        arow = (int)img * d.anchors + d.anchor_base + pix;
    }
    if (d.mode == 1) {
        float dist[4];
        #pragma unroll
        for (int side = 0; side < 4; ++side) {
            uint32_t v[16];
            tmem_ld16(t_row + (uint32_t)(side * 16), v);
            tmem_ld_wait();
            float f[16];
            #pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]) + s_bias[side * 16 + k];
            dist[side] = dfl_expectation(f);
        }
        if (arow >= 0) {
            const int y = pix / d.w, x = pix - y * d.w;
            const float4 box = dfl_box((float)x + 0.5f, (float)y + 0.5f, make_float4(dist[0], dist[1], dist[2], dist[3]), __ldg(d.stride));
            *reinterpret_cast<float4*>(d.pred + (size_t)arow * d.nout) = box;
            if (d.cand_boxes) d.cand_boxes[arow] = to_xyxy(box.x, box.y, box.z, box.w);
        }
    } else {
        float best = -INFINITY; int bi = 0x7fffffff;
        const uint32_t line = s_out + (uint32_t)row * 128u;
        const int wrow0 = row - lane;                              // first tile row of this warp
        #pragma unroll 1
        for (int cb = 0; cb < nc; cb += 32) {
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = cb + h * 16;
                if (c0 < nc) {
                    uint32_t v[16];
                    tmem_ld16(t_row + (uint32_t)c0, v);
                    tmem_ld_wait();
                    float f[16];
                    #pragma unroll
                    for (int k = 0; k < 16; ++k) f[k] = __uint_as_float(v[k]) + s_bias[c0 + k];
                    float4 r[4]; float cbest; int ci;
                    cls_chunk16(f, c0, r, cbest, ci);
                    if (c0 == 0 || cbest > best) { best = cbest; bi = ci; }   // chunks ascend: the first maximum wins (torch.max)
                    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t addr = line + (((uint32_t)(h * 4 + q) ^ (uint32_t)(row & 7)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(r[q].x), "f"(r[q].y), "f"(r[q].z), "f"(r[q].w) : "memory");
                    }
                }
            }
            __syncwarp();
            const int pieces = (nc - cb >= 32 ? 32 : nc - cb) >> 2;          // 16-byte pieces per row in this chunk
            const int piece = lane & 7;
            #pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3);                          // row of the warp's 32 this lane helps to store
                const int src = __shfl_sync(0xffffffffu, arow, rr);
                if (piece < pieces && src >= 0) {
                    const int ra = wrow0 + rr;
                    const uint32_t addr = s_out + (uint32_t)ra * 128u + (((uint32_t)piece ^ (uint32_t)(ra & 7)) << 4);
                    float4 o;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(addr));
                    *reinterpret_cast<float4*>(d.pred + (size_t)src * d.nout + 4 + cb + piece * 4) = o;
                }
            }
            __syncwarp();                                                     // the rows are rewritten by the next chunk
        }
        if (arow >= 0 && d.cand_scores) { d.cand_scores[arow] = best; d.cand_labels[arow] = bi; }
    }
}

// kMode: 0 plain (bf16 / fp32 output), 1 decode-fused epilogue, 2 up-add epilogue -- separate instantiations keep each epilogue's
// register footprint (decode state, 32 registers of prefetched partial sums) out of the plain convolution kernel
template <int kMode>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x2,
                 const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
                 const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ ConvKernelParams p) {
    constexpr bool kDec = kMode == 1, kUpAdd = kMode == 2;
    extern __shared__ unsigned char smem_dyn[];
    long long pw0 = 0, pw1 = 0, pw2 = 0;            // wait-cycle accumulators (dead code unless -DYMS_PROF)
    (void)pw0; (void)pw1; (void)pw2;
    YMS_PROF_ONLY(const long long prof_t_entry = clock64(); long long* prof = (p.prof && blockIdx.x < kNumSMs) ? p.prof + 16 * blockIdx.x : nullptr;)
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
    float* s_bias = reinterpret_cast<float*>(g_out0 + p.epi_groups * kStageOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 4 + s); };
    auto res_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 8 + s); };
    auto w_bar = [&]() { return bar0 + 8u * (2 * kMaxStages + 12); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const BiasRegs bias_regs = bias_fetch(p.bias, p.c_out);          // staged by the epilogue warps after the CTA-wide sync
    if (warp == 0) {                                  // one barrier per lane: the prologue is paid by every launch
        if (lane == 0) {
            prefetch_tmap(&tm_x); prefetch_tmap(&tm_w);
            if (p.kb2) prefetch_tmap(&tm_x2);
            prefetch_tmap(&tm_y);
            if (p.has_res) prefetch_tmap(&tm_res);
        }
        if (lane < 2 * kMaxStages + 13) {
            const bool is_tempty = lane >= 2 * kMaxStages + 4 && lane < 2 * kMaxStages + 8;
            mbar_init(bar0 + 8u * lane, is_tempty ? 4 * (p.epi_groups / p.acc_stages) : 1);
        }
        fence_barrier_init();
        __syncwarp();
        YMS_PROF_ONLY(if (prof && lane == 0) prof[15] = clock64() - prof_t_entry;)
        // weights are constants of the program: fetched before the CTA-wide sync (overlapping the TMEM allocation and the
        // bias staging) and BEFORE the grid dependency resolves, i.e. while the previous layer is still draining
        if (p.resident && elect_one()) {
            const int kpt = p.kb1 + p.kb2;
            mbar_expect_tx(w_bar(), (uint32_t)kb_total_res * (uint32_t)b_tile_bytes);
            for (int tap = 0; tap < p.taps; ++tap)
                for (int kb = 0; kb < kpt; ++kb)
                    tma_load_3d(smem_bres + (tap * kpt + kb) * b_tile_pad, &tm_w, w_bar(),
                                kb < p.kb1 ? kb * kBlockK : p.c_in1 + (kb - p.kb1) * kBlockK, 0, tap);
        }
        __syncwarp();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
        YMS_PROF_ONLY(if (prof && lane == 0) prof[14] = clock64() - prof_t_entry;)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    YMS_PROF_ONLY(if (prof && threadIdx.x == 64) prof[13] = clock64() - prof_t_entry;)
    pdl_launch_dependents();
    pdl_wait();                                   // previous grid complete: its outputs may be read, ours written

    YMS_PROF_ONLY(const long long prof_t_start = clock64();)
    const int kb_per_tap = p.kb1 + p.kb2;
    const int num_kb = p.taps * kb_per_tap;
    const int pad = p.ksize >> 1;
    const uint32_t a_bytes = (uint32_t)(p.tw * p.th) * 128u;
    const int acc_stride = p.tmem_cols / p.acc_stages;
    const uint32_t stage_tx = a_bytes + (p.resident ? 0u : (uint32_t)b_tile_bytes);

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(p, t);
                const int n0 = tc.n_tile * p.block_n;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
                    const int xin = tc.x0 * p.stride + kx - pad;
                    const int yin = tc.y0 * p.stride + ky - pad;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        mbar_wait_acc(empty_bar(stage), phase ^ 1u, pw0);
                        const uint32_t sa = smem_a0 + stage * stage_bytes;
                        const uint32_t sb = sa + kATileBytes;
                        mbar_expect_tx(full_bar(stage), stage_tx);
                        if (p.s2_dense) {
                            // dense stride-2 input viewed as (2C, W/2, 2, H/2, N): a tap selects the pixel/row parity
                            // (channel offset px*C, coordinate py) and a -1/0 shift; every box row is contiguous.
                            const int px = (kx != 1), py = (ky != 1);
                            tma_load_5d(sa, &tm_x, full_bar(stage), px * p.c_in1 + kb * kBlockK, tc.x0 - (kx == 0), py,
                                        tc.y0 - (ky == 0), tc.img);
                            if (!p.resident) tma_load_3d(sb, &tm_w, full_bar(stage), kb * kBlockK, n0, tap);
                        } else if (kb < p.kb1) {
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
            YMS_PROF_ONLY(if (prof) { prof[4] = clock64() - prof_t_start; prof[5] = pw0; })
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // ONE elected thread runs the whole loop (measured with scripts/ubench/mma_ring.cu: the per-k-block cost of the
        // issue path -- barrier wait, fence, descriptor moves, 4 x tcgen05.mma, commit -- is what bounds the small-N
        // layers, and it is ~25 % lower without the per-k-block elect/warp-sync); the try_wait of the NEXT stage is
        // issued before the MMAs of the current one so that its latency overlaps their issue.
        // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
        if (elect_one()) {
            if (p.resident) { mbar_wait_acc(w_bar(), 0u, pw2); tc_fence_after(); }
            YMS_PROF_ONLY(int ntile = 0;)
            const uint64_t hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);   // make_sw128_desc without the address
            const uint32_t a0_16 = (smem_a0 & 0x3FFFFu) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
            const uint32_t bres16 = (smem_bres & 0x3FFFFu) >> 4, btile16 = (uint32_t)b_tile_pad >> 4;
            const int tail1 = ((p.c_in1 - (p.kb1 - 1) * kBlockK) + 15) >> 4;                   // k-steps of the last block of each source
            const int tail2 = p.kb2 ? (((p.c_in2 - (p.kb2 - 1) * kBlockK) + 15) >> 4) : 4;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t ready = 0;                                      // prefetched try_wait result for (stage, phase)
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                mbar_wait_acc(tempty_bar(acc), acc_phase ^ 1u, pw1);  // epilogue has drained this accumulator
                YMS_PROF_ONLY(++ntile;)
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
                int kbi = 0;
                for (int tap = 0; tap < p.taps; ++tap) {
                    for (int kb = 0; kb < kb_per_tap; ++kb, ++kbi) {
                        if (!ready) mbar_wait_acc(full_bar(stage), phase, pw0);
                        tc_fence_after();
                        int nstage = stage + 1; uint32_t nphase = phase;
                        if (nstage == p.num_stages) { nstage = 0; nphase ^= 1u; }
                        ready = mbar_try_wait(full_bar(nstage), nphase);
                        const int ksteps = (kb == p.kb1 - 1) ? tail1 : ((kb == kb_per_tap - 1) ? tail2 : 4);
                        const uint32_t a16 = a0_16 + (uint32_t)stage * stage16;
                        const uint32_t b16 = p.resident ? bres16 + (uint32_t)kbi * btile16 : a16 + (uint32_t)(kATileBytes >> 4);
                        umma_bf16(d_tmem, hi | (uint64_t)a16, hi | (uint64_t)b16, idesc, kbi ? 1u : 0u);
                        #pragma unroll
                        for (int k = 1; k < 4; ++k)                 // +32 B (16 bf16) along K inside the swizzle atom
                            if (k < ksteps) umma_bf16(d_tmem, hi | (uint64_t)(a16 + 2 * k), hi | (uint64_t)(b16 + 2 * k), idesc, 1u);
                        umma_commit(empty_bar(stage));             // smem slot free once these MMAs retire
                        if (kbi == num_kb - 1) umma_commit(tfull_bar(acc));
                        stage = nstage; phase = nphase;
                    }
                }
                if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
            YMS_PROF_ONLY(if (prof) { prof[0] = clock64() - prof_t_start; prof[1] = pw0; prof[3] = pw1; prof[10] = pw2; prof[11] = ntile;
                                      prof[9] = prof_t_start - prof_t_entry; })
        }
        __syncwarp();
    } else {
        // ================= epilogue: up to 4 groups of 4 warps, group e drains accumulator stage e =================
        const int grp = (warp - 2) >> 2;
        const int gps = p.epi_groups / p.acc_stages;       // groups sharing one accumulator stage
        const int stage_id = grp / gps, sub_id = grp - stage_id * gps;
        bias_stage(s_bias, bias_regs, p.bias, p.c_out, p.bias_pad, p.act);
        {
            EpiShared e;
            e.tm_y = &tm_y; e.tm_res = &tm_res;
            e.res_bar = res_bar(grp);
            e.s_out = smem_out0 + grp * kStageOutBytes;
            e.s_bias = s_bias;
            e.block_n = p.block_n; e.c_out = p.c_out; e.act = p.act; e.has_res = p.has_res;
            e.out_bytes = a_bytes;
            e.bar_id = 1 + grp;
            e.leader = ((warp - 2) & 3) == 0 && lane == 0;
            e.row = (warp & 3) * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(stage_id * acc_stride);
            const int n_chunks = (p.block_n + 63) >> 6;
            uint32_t res_phase = 0u, acc_phase = 0u;
            // up-add: this thread's row of the half-resolution partial sums for tile t (nearest upsample: (y, x) reads (y/2, x/2)),
            // offset to the tile's first output channel
            auto up_row_of = [&](int t) -> const float* {
                const TileCoord c = decode_tile(p, t);
                uint32_t m = (uint32_t)(c.x0 + e.row);
                if (m >= (uint32_t)p.out_w) m = (uint32_t)p.out_w - 1u;                 // rows past the last pixel are clipped by the store
                const uint32_t img = m / (uint32_t)p.up_hw, rem = m - img * (uint32_t)p.up_hw;
                const uint32_t y = rem / (uint32_t)p.up_w, x = rem - y * (uint32_t)p.up_w;
                return p.up + ((size_t)img * (size_t)(p.up_hw >> 2) + (size_t)(y >> 1) * (size_t)(p.up_w >> 1) + (x >> 1)) * (size_t)p.up_ps
                       + c.n_tile * p.block_n;
            };
            auto up_groups = [&](int ch) { const int g = (p.block_n - ch * 64) >> 4; return g < 4 ? g : 4; };
            const int t_first = blockIdx.x + stage_id * gridDim.x, t_step = p.acc_stages * gridDim.x;
            UpRegs upr;
            const float* up_cur = nullptr;
            if (kUpAdd && t_first < p.total_tiles && sub_id < n_chunks) {
                up_cur = up_row_of(t_first);
                up_regs_load(upr, up_cur + sub_id * 64, e.row, up_groups(sub_id));
            }
            for (int t = t_first; t < p.total_tiles; t += t_step) {
                const TileCoord tc = decode_tile(p, t);
                EpiTile tl; tl.n0 = tc.n_tile * p.block_n; tl.x0 = tc.x0; tl.y0 = tc.y0; tl.img = tc.img;
                const float* up_next = (kUpAdd && t + t_step < p.total_tiles) ? up_row_of(t + t_step) : nullptr;
                mbar_wait_acc(tfull_bar(stage_id), acc_phase, pw0);
                acc_phase ^= 1u;
                tc_fence_after();
                if constexpr (kDec) {
                    epilogue_decode(p.dec, s_bias, p.c_out, t_row, e.s_out, e.row, lane, (uint32_t)(tl.x0 + e.row), (uint32_t)p.out_w);
                } else if constexpr (kUpAdd) {
                    for (int ch = sub_id; ch < n_chunks; ch += gps) {
                        const bool more = ch + gps < n_chunks;                          // next unit: this tile's next chunk, else the next tile's first
                        const float* nx = more ? up_cur + (ch + gps) * 64 : (up_next ? up_next + sub_id * 64 : nullptr);
                        epilogue_chunk_bf16<true>(e, res_phase, t_row, tl, ch, &upr, nx, nx ? up_groups(more ? ch + gps : sub_id) : 0);
                    }
                    up_cur = up_next;
                } else if (p.out_f32) {
                    const int n_chunks32 = (p.block_n + 31) >> 5;
                    for (int ch = sub_id; ch < n_chunks32; ch += gps) epilogue_chunk_f32(e, t_row, tl, ch);
                } else {
                    for (int ch = sub_id; ch < n_chunks; ch += gps) epilogue_chunk_bf16(e, res_phase, t_row, tl, ch);
                }
                tc_fence_before();                               // all TMEM reads of this stage by this warp are done
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(stage_id));
            }
            if (e.leader) tma_store_wait_read<0>();
            YMS_PROF_ONLY(if (prof && warp == 2 && lane == 0) { prof[7] = clock64() - prof_t_start; prof[8] = pw0; })
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
    YMS_PROF_ONLY(if (prof && threadIdx.x == 0) prof[12] = clock64() - prof_t_entry;)
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (yms_conv_params.variant == 5; bf16 output, no decode / up-add epilogue).  Two CTAs of a cluster on the
// two SMs of a TPC run ONE tcgen05.mma of M = 256 per k-step: CTA r owns M tile 2c + r of cluster item c (same N tile),
// loads its own A tile and HALF of the weight tile (N/2 rows); the accumulator rows of its tile stay in its own TMEM, so the
// epilogue is the single-CTA one.  What it buys on the layers this kernel runs (1x1 with K >= 256 on the 40x40 / 20x20 maps,
// 3x3/s2, 3x3 on 20x20): per k-block the single-thread issue path costs ~290 + 4 x 36 cycles against 4 x 64 cycles of
// tensor-pipe work at N = 128 -- issue-bound -- and every CTA re-streams the whole weight tile from L2; the pair issues the
// same number of instructions for twice the work and fetches each weight byte once per TPC.
// Protocol as in conv3x3_pair_kernel: full barriers in the leader (rank 0), signalled by the cta_group::2 TMA loads of both
// CTAs; commits multicast to the empty / accumulator-full barriers of both; both epilogues arrive on the leader's
// accumulator-empty barrier.
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x2,
                      const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
                      const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ ConvKernelParams p) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;          // same offset in both CTAs (same kernel, same dynamic size)
    unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t rank = cluster_ctarank();
    const bool leader_cta = rank == 0;
    const int half_n = p.block_n >> 1;
    const int b_tile_bytes = half_n * 128;
    const int b_tile_pad = (b_tile_bytes + 1023) & ~1023;
    const int stage_bytes = kATileBytes + (p.resident ? 0 : b_tile_pad);
    const int kb_total_res = p.taps * (p.kb1 + p.kb2);
    const int ring_bytes = p.num_stages * stage_bytes + (p.resident ? kb_total_res * b_tile_pad : 0);
    const uint32_t smem_a0 = base;
    const uint32_t smem_bres = base + p.num_stages * stage_bytes;
    const uint32_t smem_out0 = base + ring_bytes;
    unsigned char* g_out0 = gbase + ring_bytes;
    float* s_bias = reinterpret_cast<float*>(g_out0 + p.epi_groups * kStageOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 4 + s); };
    auto res_bar = [&](int s) { return bar0 + 8u * (2 * kMaxStages + 8 + s); };
    auto w_bar = [&]() { return bar0 + 8u * (2 * kMaxStages + 12); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 13);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int gps = p.epi_groups / p.acc_stages;

    const BiasRegs bias_regs = bias_fetch(p.bias, p.c_out);          // staged by the epilogue warps after the CTA-wide sync
    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tm_x); prefetch_tmap(&tm_w);
            if (p.kb2) prefetch_tmap(&tm_x2);
            prefetch_tmap(&tm_y);
            if (p.has_res) prefetch_tmap(&tm_res);
        }
        if (lane < 2 * kMaxStages + 13) {
            const bool is_tempty = lane >= 2 * kMaxStages + 4 && lane < 2 * kMaxStages + 8;
            mbar_init(bar0 + 8u * lane, is_tempty ? 2 * 4 * gps : 1);   // accumulator-empty (leader's): epilogue warps of BOTH CTAs
        }
        fence_barrier_init();
        __syncwarp();
    }
    cluster_sync_all();                                    // the peer's barriers exist before anything signals them
    if (warp == 0 && p.resident && elect_one()) {
        const int kpt = p.kb1 + p.kb2;
        if (leader_cta) mbar_expect_tx(w_bar(), 2u * (uint32_t)kb_total_res * (uint32_t)b_tile_bytes);
        for (int tap = 0; tap < p.taps; ++tap)
            for (int kb = 0; kb < kpt; ++kb)
                tma_load_3d_2sm(smem_bres + (tap * kpt + kb) * b_tile_pad, &tm_w, w_bar(),
                                kb < p.kb1 ? kb * kBlockK : p.c_in1 + (kb - p.kb1) * kBlockK, (int)rank * half_n, tap);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    const int kb_per_tap = p.kb1 + p.kb2;
    const int num_kb = p.taps * kb_per_tap;
    const int pad = p.ksize >> 1;
    const uint32_t a_bytes = (uint32_t)(p.tw * p.th) * 128u;
    const int acc_stride = p.tmem_cols / p.acc_stages;
    const uint32_t stage_tx = 2u * (a_bytes + (p.resident ? 0u : (uint32_t)b_tile_bytes));
    const int cid = (int)cluster_id_x(), ncl = (int)num_clusters_x();
    // cluster item u = (N tile, pair of M tiles); this CTA's tile as an index of the single-CTA enumeration (N tile fastest)
    auto my_tile = [&](int u, bool& valid) {
        const int mp = (int)fast_div((uint32_t)u, p.mg_n_tiles), nt = u - mp * p.n_tiles;
        int m = 2 * mp + (int)rank;
        valid = m < p.m_tiles;
        if (!valid) m = p.m_tiles - 1;                     // odd tile count: the spare CTA re-loads its neighbour's tile and stores nothing
        return m * p.n_tiles + nt;
    };

    if (warp == 0) {
        // ================= TMA producer (both CTAs) =================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int u = cid; u < p.total_tiles; u += ncl) {
                bool valid;
                const TileCoord tc = decode_tile(p, my_tile(u, valid));
                const int n0 = tc.n_tile * p.block_n + (int)rank * half_n;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
                    const int xin = tc.x0 * p.stride + kx - pad;
                    const int yin = tc.y0 * p.stride + ky - pad;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t sa = smem_a0 + stage * stage_bytes;
                        const uint32_t sb = sa + kATileBytes;
                        if (leader_cta) mbar_expect_tx(full_bar(stage), stage_tx);
                        if (p.s2_dense) {
                            const int px = (kx != 1), py = (ky != 1);
                            asm volatile(
                                "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                                ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tm_x)), "r"(full_bar(stage) & kPeerBitMask), "r"(px * p.c_in1 + kb * kBlockK),
                                  "r"(tc.x0 - (kx == 0)), "r"(py), "r"(tc.y0 - (ky == 0)), "r"(tc.img) : "memory");
                            if (!p.resident) tma_load_3d_2sm(sb, &tm_w, full_bar(stage), kb * kBlockK, n0, tap);
                        } else if (kb < p.kb1) {
                            tma_load_4d_2sm(sa, &tm_x, full_bar(stage), kb * kBlockK, xin, yin, tc.img);
                            if (!p.resident) tma_load_3d_2sm(sb, &tm_w, full_bar(stage), kb * kBlockK, n0, tap);
                        } else {
                            tma_load_4d_2sm(sa, &tm_x2, full_bar(stage), (kb - p.kb1) * kBlockK, xin, yin, tc.img);
                            if (!p.resident) tma_load_3d_2sm(sb, &tm_w, full_bar(stage), p.c_in1 + (kb - p.kb1) * kBlockK, n0, tap);
                        }
                        if (++stage == p.num_stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: one elected thread of the LEADER =================
        if (leader_cta && elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            if (p.resident) { mbar_wait(w_bar(), 0u); tc_fence_after(); }
            const uint64_t hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t a0_16 = (smem_a0 & 0x3FFFFu) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
            const uint32_t bres16 = (smem_bres & 0x3FFFFu) >> 4, btile16 = (uint32_t)b_tile_pad >> 4;
            const int tail1 = ((p.c_in1 - (p.kb1 - 1) * kBlockK) + 15) >> 4;
            const int tail2 = p.kb2 ? (((p.c_in2 - (p.kb2 - 1) * kBlockK) + 15) >> 4) : 4;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t ready = 0;
            for (int u = cid; u < p.total_tiles; u += ncl) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
                int kbi = 0;
                for (int tap = 0; tap < p.taps; ++tap) {
                    for (int kb = 0; kb < kb_per_tap; ++kb, ++kbi) {
                        if (!ready) mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        int nstage = stage + 1; uint32_t nphase = phase;
                        if (nstage == p.num_stages) { nstage = 0; nphase ^= 1u; }
                        ready = mbar_try_wait(full_bar(nstage), nphase);
                        const int ksteps = (kb == p.kb1 - 1) ? tail1 : ((kb == kb_per_tap - 1) ? tail2 : 4);
                        const uint32_t a16 = a0_16 + (uint32_t)stage * stage16;
                        const uint32_t b16 = p.resident ? bres16 + (uint32_t)kbi * btile16 : a16 + (uint32_t)(kATileBytes >> 4);
                        umma_bf16_2sm(d_tmem, hi | (uint64_t)a16, hi | (uint64_t)b16, idesc, kbi ? 1u : 0u);
                        #pragma unroll
                        for (int k = 1; k < 4; ++k)
                            if (k < ksteps) umma_bf16_2sm(d_tmem, hi | (uint64_t)(a16 + 2 * k), hi | (uint64_t)(b16 + 2 * k), idesc, 1u);
                        umma_commit_2sm(empty_bar(stage));
                        if (kbi == num_kb - 1) umma_commit_2sm(tfull_bar(acc));
                        stage = nstage; phase = nphase;
                    }
                }
                if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue (both CTAs, each its own tile) =================
        const int grp = (warp - 2) >> 2;
        const int stage_id = grp / gps, sub_id = grp - stage_id * gps;
        bias_stage(s_bias, bias_regs, p.bias, p.c_out, p.bias_pad, p.act);
        EpiShared e;
        e.tm_y = &tm_y; e.tm_res = &tm_res;
        e.res_bar = res_bar(grp);
        e.s_out = smem_out0 + grp * kStageOutBytes;
        e.s_bias = s_bias;
        e.block_n = p.block_n; e.c_out = p.c_out; e.act = p.act; e.has_res = p.has_res;
        e.out_bytes = a_bytes;
        e.bar_id = 1 + grp;
        e.leader = ((warp - 2) & 3) == 0 && lane == 0;
        e.row = (warp & 3) * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(stage_id * acc_stride);
        const int n_chunks = (p.block_n + 63) >> 6;
        uint32_t res_phase = 0u, acc_phase = 0u;
        for (int u = cid + stage_id * ncl; u < p.total_tiles; u += p.acc_stages * ncl) {
            bool valid;
            const TileCoord tc = decode_tile(p, my_tile(u, valid));
            EpiTile tl; tl.n0 = tc.n_tile * p.block_n; tl.x0 = tc.x0; tl.y0 = tc.y0; tl.img = tc.img;
            mbar_wait(tfull_bar(stage_id), acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            if (valid)
                for (int ch = sub_id; ch < n_chunks; ch += gps) epilogue_chunk_bf16(e, res_phase, t_row, tl, ch);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(stage_id), 0u);
        }
        if (e.leader) tma_store_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, (uint32_t)p.tmem_cols);
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
               int box_x, int box_y, int estride, const char* what, bool f32) {
    uint64_t dims[4]; uint64_t strides[3]; uint32_t box[4]; uint32_t es[4] = {1, (uint32_t)estride, (uint32_t)estride, 1};
    dims[0] = (uint64_t)c;
    strides[0] = (uint64_t)ps * (f32 ? 4 : 2);
    if (flat) {
        dims[1] = (uint64_t)batch * h * w; dims[2] = 1; dims[3] = 1;
        strides[1] = strides[0] * dims[1]; strides[2] = strides[1];
    } else {
        dims[1] = (uint64_t)w; dims[2] = (uint64_t)h; dims[3] = (uint64_t)batch;
        strides[1] = strides[0] * (uint64_t)w; strides[2] = strides[1] * (uint64_t)h;
    }
    box[0] = f32 ? 32 : kBlockK;            // 128-byte rows either way
    box[1] = (uint32_t)box_x; box[2] = (uint32_t)box_y; box[3] = 1;
    return encode_map(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, what);
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
    if (q->out_dtype == YMS_DTYPE_F32 && ((q->y_pixel_stride % 4) || (q->c_out % 4))) return fail(YMS_E_ARG, "conv: f32 output needs c_out, pixel stride % 4");
    if (q->c_in2 && (!q->x2 || !al16(q->x2) || (q->x2_pixel_stride % 8) || q->x2_pixel_stride < q->c_in2))
        return fail(YMS_E_ARG, "conv: bad second source");
    if (q->residual && (!al16(q->residual) || (q->res_pixel_stride % 8) || q->res_pixel_stride < q->c_out))
        return fail(YMS_E_ARG, "conv: bad residual");

    yms_conv_plan* pl = new (std::nothrow) yms_conv_plan();
    if (!pl) return fail(YMS_E_ARG, "conv: out of host memory");
    pl->kind = 0;
    if (q->variant < 0 || q->variant > 7) { delete pl; return fail(YMS_E_ARG, "conv: variant must be 0..7"); }
    if (q->variant == 4) {                                       // stride-2 pair-line kernel, pair-packed weights (conv3x3.cu)
        int rc4 = conv3_s2pair_plan_init(pl, q);
        if (rc4) { delete pl; return rc4; }
        *out = pl;
        return 0;
    }
    if (q->ksize == 3 && q->stride == 1 && q->out_dtype == YMS_DTYPE_BF16 && q->c_in2 == 0 && q->variant != 1 && q->variant != 6) {
        int rc3 = conv3_plan_init(pl, q);
        if (rc3) { delete pl; return rc3; }
        *out = pl;
        return 0;
    }
    if (q->variant == 7) { delete pl; return fail(YMS_E_UNSUPPORTED, "conv (variant 7): 3x3 stride-1 convolutions with bf16 output only"); }
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
        // spill into the next tile's channels); among the widths that pad c_out the least, the one with the shortest makespan
        // over the persistent CTAs / clusters: waves x (tile width + a fixed per-tile cost of ~32 columns).  A 20 x 20 map at
        // batch 32 has 100 M tiles: 512 channels as 2 x 256 are 200 tiles = 2 waves of 256 columns on 148 CTAs, as 4 x 128 they
        // are 3 waves of 128.
        int best_pad = 1 << 30;
        for (int bn = 64; bn <= 256; bn += 64) {
            const int padded = ceil_div(q->c_out, bn) * bn;
            if (padded < best_pad) best_pad = padded;
        }
        const int m_tiles0 = kp.tiles_x * kp.tiles_y * kp.batch;
        const bool pair0 = (q->variant == 5 || q->variant == 6);
        const long long m_items = pair0 ? ceil_div(m_tiles0, 2) : m_tiles0, ctas = pair0 ? kNumSMs / 2 : kNumSMs;
        long long best_cost = -1;
        for (int bn = 64; bn <= 256; bn += 64) {
            if (ceil_div(q->c_out, bn) * bn != best_pad) continue;
            const long long items = m_items * ceil_div(q->c_out, bn);
            const long long cost = ((items + ctas - 1) / ctas) * (bn + 32);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; kp.block_n = bn; }
        }
        kp.n_tiles = ceil_div(q->c_out, kp.block_n);
    }
    kp.c_out = q->c_out;
    kp.c_in1 = q->c_in; kp.c_in2 = q->c_in2;
    kp.kb1 = ceil_div(q->c_in, kBlockK); kp.kb2 = ceil_div(q->c_in2, kBlockK);
    kp.ksize = q->ksize; kp.taps = q->ksize * q->ksize; kp.stride = q->stride;
    kp.act = q->act ? 1 : 0; kp.out_f32 = (q->out_dtype == YMS_DTYPE_F32); kp.has_res = q->residual ? 1 : 0;
    kp.total_tiles = kp.tiles_x * kp.tiles_y * kp.batch * kp.n_tiles;
    // variant 5 (outside the 3x3/s1 halo kernels, which have their own pair variant): CTA-pair kernel, bf16 output only
    kp.pair = (q->variant == 5 || q->variant == 6) ? 1 : 0;
    kp.m_tiles = kp.tiles_x * kp.tiles_y * kp.batch;
    if (kp.pair) {
        if (q->out_dtype != YMS_DTYPE_BF16 || kp.m_tiles < 2) { delete pl; return fail(YMS_E_UNSUPPORTED, "conv (variant 5): needs bf16 output and at least two M tiles"); }
        kp.total_tiles = ceil_div(kp.m_tiles, 2) * kp.n_tiles;                    // cluster items
    }
    kp.bias_pad = kp.n_tiles * kp.block_n + 64;
    kp.bias = q->bias;
    kp.y_f32 = kp.out_f32 ? reinterpret_cast<float*>(q->y) : nullptr;
    kp.y_ps = q->y_pixel_stride;

    const int b_bytes = ((kp.pair ? kp.block_n / 2 : kp.block_n) * 128 + 1023) & ~1023;     // pair: half a weight tile per CTA
    const int res_bytes = kp.taps * (kp.kb1 + kp.kb2) * b_bytes;
    // "half" CTAs (yms_debug_set_option("conv_half", 1), experimental): 2 epilogue groups, 2 x 128 TMEM columns, <= 113 KB -> two CTAs per SM.  Only for
    // N <= 128 (two accumulator stages must remain) and when at least 3 ring stages fit beside resident / streamed weights.
    const bool want_half = g_opt.conv_half != 0 && !kp.pair;
    int groups = kEpiGroups, limit = kSmemLimit;
    if (want_half && kp.block_n <= 128) {
        const int fixed_h = 2 * kStageOutBytes + kp.bias_pad * 4 + (2 * kMaxStages + 16) * 8 + 1024;
        const bool res_h = kp.n_tiles == 1 && kSmemLimitHalf - fixed_h - res_bytes >= 3 * kATileBytes;
        if (res_h || (kSmemLimitHalf - fixed_h) / (kATileBytes + b_bytes) >= 3) { groups = 2; limit = kSmemLimitHalf; }
    }
    kp.epi_groups = groups; kp.tmem_cols = 128 * groups;
    const int fixed = groups * kStageOutBytes + kp.bias_pad * 4 + (2 * kMaxStages + 16) * 8 + 1024 /* alignment slack */;
    kp.acc_stages = groups == 2 ? 2 : (kp.block_n <= 128 ? 4 : 2);
    kp.mg_n_tiles = fast_div_magic(kp.n_tiles); kp.mg_tiles_x = fast_div_magic(kp.tiles_x); kp.mg_tiles_y = fast_div_magic(kp.tiles_y);
    // small weight sets stay resident for the whole persistent CTA (no per-tile re-fetch from L2)
    kp.resident = (kp.n_tiles == 1 && limit - fixed - res_bytes >= (groups == 2 ? 3 : 4) * kATileBytes) ? 1 : 0;
    const int stage_bytes = kATileBytes + (kp.resident ? 0 : b_bytes);
    int stages = (limit - fixed - (kp.resident ? res_bytes : 0)) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) { delete pl; return fail(YMS_E_UNSUPPORTED, "conv: tile does not fit in shared memory"); }
    kp.num_stages = stages;
    pl->smem = (size_t)stages * stage_bytes + (kp.resident ? res_bytes : 0) + fixed;
    const int max_ctas = kNumSMs * (groups == 2 ? 2 : 1);
    pl->grid = kp.total_tiles < max_ctas ? kp.total_tiles : max_ctas;
    if (kp.pair) pl->grid = 2 * (kp.total_tiles < kNumSMs / 2 ? kp.total_tiles : kNumSMs / 2);      // clusters of two CTAs
    pl->threads = 64 + groups * kEpiGroupThreads;

    int rc;
    const int K_total = q->c_in + q->c_in2;
    const int bx = kp.tw * q->stride, by = kp.th * q->stride;
    kp.s2_dense = (q->stride == 2 && q->x_pixel_stride == q->c_in && q->c_in2 == 0) ? 1 : 0;
    if (kp.s2_dense) {
        const uint64_t C = (uint64_t)q->c_in, W = (uint64_t)q->in_w, H = (uint64_t)q->in_h;
        uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, (uint64_t)q->batch};
        uint64_t strides[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
        uint32_t box[5] = {kBlockK, (uint32_t)kp.tw, 1, (uint32_t)kp.th, 1};
        uint32_t es[5] = {1, 1, 1, 1, 1};
        if ((rc = encode_map(&pl->tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, q->x, dims, strides, box, es, "x(s2 dense)"))) { delete pl; return rc; }
    } else
    if ((rc = encode_act(&pl->tm_x, q->x, q->c_in, q->x_pixel_stride, q->batch, q->in_h, q->in_w, flat, bx, by, q->stride, "x"))) { delete pl; return rc; }
    if (q->c_in2) {
        if ((rc = encode_act(&pl->tm_x2, q->x2, q->c_in2, q->x2_pixel_stride, q->batch, q->in_h, q->in_w, flat, bx, by, q->stride, "x2"))) { delete pl; return rc; }
    } else pl->tm_x2 = pl->tm_x;
    {
        uint64_t dims[3] = {(uint64_t)K_total, (uint64_t)q->c_out, (uint64_t)kp.taps};
        uint64_t strides[2] = {(uint64_t)K_total * 2, (uint64_t)K_total * 2 * (uint64_t)q->c_out};
        uint32_t box[3] = {kBlockK, (uint32_t)(kp.pair ? kp.block_n / 2 : kp.block_n), 1};
        uint32_t es[3] = {1, 1, 1};
        if ((rc = encode_map(&pl->tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->weight, dims, strides, box, es, "w"))) { delete pl; return rc; }
    }
    if ((rc = encode_act(&pl->tm_y, q->y, q->c_out, q->y_pixel_stride, q->batch, out_h, out_w, flat, kp.tw, kp.th, 1, "y", kp.out_f32 != 0))) { delete pl; return rc; }
    if (q->residual) {
        if ((rc = encode_act(&pl->tm_res, q->residual, q->c_out, q->res_pixel_stride, q->batch, out_h, out_w, flat, kp.tw, kp.th, 1, "res"))) { delete pl; return rc; }
    } else pl->tm_res = pl->tm_x;

    const double m = (double)q->batch * out_h * out_w;
    pl->flops = 2.0 * m * q->c_out * (double)K_total * kp.taps;
    pl->bytes = 2.0 * (double)q->batch * q->in_h * q->in_w * K_total + (kp.out_f32 ? 4.0 : 2.0) * m * q->c_out +
                2.0 * (double)kp.taps * q->c_out * K_total + (q->residual ? 2.0 * m * q->c_out : 0.0);

    static std::atomic<unsigned long long> attr_seen{0};
    if (first_use_on_device(attr_seen)) {
        cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (e != cudaSuccess) { delete pl; return fail((int)e, "conv: smem attribute: %s", cudaGetErrorString(e)); }
    }
    *out = pl;
    return 0;
}

extern "C" int yms_conv_plan_fuse_decode(yms_conv_plan* pl, const yms_decode_fusion* f) {
    if (!pl || !f) return fail(YMS_E_ARG, "fuse_decode: null argument");
    const ConvKernelParams& kp = pl->kp;
    if (pl->kind != 0 || kp.ksize != 1 || kp.act || !kp.out_f32 || kp.has_res || kp.n_tiles != 1)
        return fail(YMS_E_UNSUPPORTED, "fuse_decode: the plan must be a linear 1x1 convolution with f32 output");
    if (f->branch != 1 && f->branch != 2) return fail(YMS_E_ARG, "fuse_decode: branch must be 1 (box) or 2 (class)");
    const int nc = f->num_classes;
    if (nc <= 0 || (nc % 16) || nc > 128) return fail(YMS_E_UNSUPPORTED, "fuse_decode: num_classes must be a multiple of 16, <= 128");
    if (kp.c_out != (f->branch == 1 ? 4 * kRegMax : nc) || kp.block_n != kp.c_out || kp.acc_stages != kp.epi_groups)
        return fail(YMS_E_ARG, "fuse_decode: c_out must be 64 (box branch) or num_classes (class branch)");
    const long long hw = (long long)f->map_h * f->map_w;
    if (f->map_h <= 0 || f->map_w <= 0 || (long long)kp.out_w % hw) return fail(YMS_E_ARG, "fuse_decode: map size does not divide the plan's pixels");
    if (f->anchor_base < 0 || (long long)f->anchor_base + hw > f->anchors) return fail(YMS_E_ARG, "fuse_decode: anchor range");
    if (!f->stride || !f->pred || ((uintptr_t)f->pred & 15)) return fail(YMS_E_ARG, "fuse_decode: stride / pred (16-byte aligned) required");
    if (f->branch == 1 && f->cand_boxes && ((uintptr_t)f->cand_boxes & 15)) return fail(YMS_E_ARG, "fuse_decode: cand_boxes must be 16-byte aligned");
    if (f->branch == 2 && ((f->cand_scores == nullptr) != (f->cand_labels == nullptr)))
        return fail(YMS_E_ARG, "fuse_decode: candidate scores and labels must come together");
    DecodeFuse& d = pl->kp.dec;
    d.mode = f->branch; d.hw = (int)hw; d.w = f->map_w; d.anchor_base = f->anchor_base; d.anchors = f->anchors; d.nout = 4 + nc;
    d.stride = f->stride; d.pred = f->pred;
    d.cand_boxes = f->branch == 1 ? reinterpret_cast<float4*>(f->cand_boxes) : nullptr;
    d.cand_scores = f->branch == 2 ? f->cand_scores : nullptr;
    d.cand_labels = f->branch == 2 ? f->cand_labels : nullptr;
    // algorithmic bytes: the f32 logits are no longer written; the decoded rows are
    const double m = (double)kp.out_w;
    pl->bytes += m * ((f->branch == 1 ? 16.0 + (f->cand_boxes ? 16.0 : 0.0) : 4.0 * nc + (f->cand_scores ? 8.0 : 0.0)) - 4.0 * kp.c_out);
    return 0;
}

extern "C" int yms_conv_plan_add_upsampled(yms_conv_plan* pl, const float* t, int64_t t_pixel_stride, int out_h, int out_w) {
    if (!pl || !t) return fail(YMS_E_ARG, "add_upsampled: null argument");
    ConvKernelParams& kp = pl->kp;
    if (pl->kind != 0 || kp.ksize != 1 || kp.out_f32 || kp.dec.mode || kp.pair)
        return fail(YMS_E_UNSUPPORTED, "add_upsampled: the plan must be a 1x1 convolution with bf16 output (not the CTA-pair variant)");
    if (out_h <= 0 || out_w <= 0 || ((out_h | out_w) & 1) || (long long)kp.out_w % ((long long)out_h * out_w))
        return fail(YMS_E_ARG, "add_upsampled: even H, W that divide the plan's pixels are required");
    if ((kp.c_out % 16) || kp.n_tiles * kp.block_n != kp.c_out)      // every accumulator column must be a real channel of t
        return fail(YMS_E_UNSUPPORTED, "add_upsampled: c_out must be a multiple of 16 (of 64 beyond 256 channels)");
    if (t_pixel_stride < kp.c_out || (t_pixel_stride % 4) || ((uintptr_t)t & 15))
        return fail(YMS_E_ARG, "add_upsampled: pixel stride %% 4 >= c_out and a 16-byte aligned tensor are required");
    kp.up = t; kp.up_ps = t_pixel_stride; kp.up_w = out_w; kp.up_hw = out_h * out_w;
    pl->bytes += 4.0 * (double)(kp.out_w / 4) * kp.c_out;                          // the partial sums are read once
    return 0;
}

extern "C" int yms_conv_plan_run(const yms_conv_plan* pl, void* stream) {
    if (!pl) return fail(YMS_E_ARG, "conv: null plan");
    if (pl->kind == 1) return conv3_plan_run(pl, (cudaStream_t)stream);
    ConvKernelParams kp = pl->kp;
    kp.prof = g_prof_buf;
    cudaError_t le = kp.pair
        ? launch_pdl_cluster(conv_gemm_pair_kernel, pl->grid, pl->threads, pl->smem, (cudaStream_t)stream, 2, pl->tm_x, pl->tm_x2, pl->tm_w, pl->tm_y, pl->tm_res, kp)
        : launch_pdl(kp.dec.mode ? conv_gemm_kernel<1> : kp.up ? conv_gemm_kernel<2> : conv_gemm_kernel<0>, pl->grid, pl->threads, pl->smem, (cudaStream_t)stream, pl->tm_x, pl->tm_x2, pl->tm_w,
                     pl->tm_y, pl->tm_res, kp);
    if (le != cudaSuccess) return fail((int)le, "conv_gemm_kernel launch: %s", cudaGetErrorString(le));
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
