// Shared epilogue of the tcgen05 convolution kernels.
//
// The epilogue of one tile is a latency chain (wait accumulator -> tcgen05.ld -> bias/SiLU ->
// swizzled smem -> fence -> TMA store); ncu showed the MMA/TMA side idle behind it on the small-N,
// high-resolution layers.  It is therefore run by up to kEpiGroups independent groups of 4 warps
// (one warp per TMEM lane quadrant), group e working on accumulator stage e with its own staging
// buffer, named barrier and TMA-store bulk groups, so that several tiles drain concurrently.
#pragma once
#include "tc_ptx.cuh"

namespace yms {
namespace tc {

constexpr int kEpiGroups = 4;
constexpr int kEpiGroupThreads = 128;
constexpr int kConvThreads = 64 + kEpiGroups * kEpiGroupThreads;      // producer warp + MMA warp + 16 epilogue warps
constexpr int kStageOutBytes = 128 * 128;                             // one 128-row x 64-channel bf16 staging tile

// vh (virtual-row tiling, see conv3x3_pair_kernel<true>): images are `vh` virtual rows apart, y0 is the tile's first virtual row
// inside image `img`; a tile whose 16 rows run past vh continues in image img + 1 at row y0 - vh.
// kStore == 2 (the stem's half tiles): rows 0..63 are stored at (x0, y0, img), rows 64..127 at (x1, y1, img1).
struct EpiTile { int n0, x0, y0, img; int vh = 0; int x1 = 0, y1 = 0, img1 = 0; };

__device__ __forceinline__ unsigned long long epi_pack(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void epi_unpack(unsigned long long v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long epi_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Bias staging, split around the CTA-wide sync of the prologue.  The epilogue threads (threadIdx.x >= 64 in every conv kernel)
// fetch the bias into registers BEFORE that sync (bias_fetch) and stage it in shared memory after it, among themselves
// (bias_stage, in the epilogue branch).  Staged before the sync, its DRAM round trip -- a few hundred bytes nobody touched since
// the previous step -- was the longest leg of a prologue that every launch pays and that, at one CTA per SM, nothing overlaps.
struct BiasRegs { float r0, r1; };
__device__ __forceinline__ BiasRegs bias_fetch(const float* bias, int c_out) {
    const int et = (int)threadIdx.x - 64, n_et = (int)blockDim.x - 64;
    BiasRegs b{0.f, 0.f};
    if (et >= 0) {
        if (et < c_out) b.r0 = bias[et];
        if (et + n_et < c_out) b.r1 = bias[et + n_et];
    }
    return b;
}
// act: the SiLU epilogue works on h = x / 2, so 0.5 * bias is what it adds
__device__ __forceinline__ void bias_stage(float* s_bias, const BiasRegs& b, const float* bias, int c_out, int bias_pad, int act) {
    const int et = (int)threadIdx.x - 64, n_et = (int)blockDim.x - 64;
    if (et < bias_pad) s_bias[et] = act ? 0.5f * b.r0 : b.r0;
    if (et + n_et < bias_pad) s_bias[et + n_et] = act ? 0.5f * b.r1 : b.r1;
    for (int i = et + 2 * n_et; i < bias_pad; i += n_et) s_bias[i] = (i < c_out) ? (act ? 0.5f * bias[i] : bias[i]) : 0.f;
    epi_all_bar_sync(n_et);                                // visible to every epilogue warp
}

struct EpiShared {
    const CUtensorMap* tm_y; const CUtensorMap* tm_res;
    uint32_t res_bar;          // mbarrier of this group for the residual TMA load
    uint32_t s_out;            // this group's staging buffer (1024-aligned)
    const float* s_bias;       // bias (act: 0.5 * bias) in shared memory, indexed by output channel
    int block_n, c_out, act, has_res;
    uint32_t out_bytes;        // bytes of one residual / store box
    int bar_id;                // named barrier of the group
    bool leader;               // thread that issues the group's TMA traffic
    int row;                   // accumulator lane == tile row of this thread
    const CUtensorMap* tm_res_row = nullptr;   // kVy: residual / output maps with a two-row (2 x 8 pixel) box
    const CUtensorMap* tm_y_row = nullptr;
};

// bf16 output of ONE 64-channel chunk through swizzled staging + TMA store (+ residual TMA-loaded
// into the same buffer).  Callers deal (sub-tile, chunk) units to the groups that share a stage.
// kUp: `up_row` points at this thread's pixel of an fp32 [.., c_out] tensor of PARTIAL SUMS that is added to the accumulator before
// bias / activation (the low-resolution half of a 1x1 convolution over cat[upsample2x(a), b], see yms_conv_plan_add_upsampled).
// The partial sums of a chunk live in UpRegs (this lane's HALF of its 16-channel groups, see below); the caller fills them for the
// first chunk (up_regs_load) and each call refills a group's registers from `next` -- the same lane's values of the NEXT chunk
// this thread will process, `next_groups` 16-channel groups of it -- as soon as the group is consumed: a whole chunk of work
// (>= an L2 round trip) lies between a load and its use.  A one-group look-ahead left most of that latency exposed four times per
// chunk and made the up-add cost as much as the GEMM.
struct UpRegs { float4 u[8]; };
__device__ __forceinline__ void up_regs_load(UpRegs& r, const float* chunk_row, int row, int groups) {
    const float4* q = reinterpret_cast<const float4*>(chunk_row) + (row & 1) * 2;
    #pragma unroll
    for (int g = 0; g < 4; ++g)
        if (g < groups) { r.u[2 * g] = __ldg(q + 4 * g); r.u[2 * g + 1] = __ldg(q + 4 * g + 1); }
}
// kStore: 0 one box per tile, 1 virtual-row tiling (conv3x3_pair_kernel<true>), 2 two half-tile boxes (stem_tma_kernel)
template <bool kUp = false, int kStore = 0>
__device__ __forceinline__ void epilogue_chunk_bf16(const EpiShared& e, uint32_t& res_phase, uint32_t t_row, const EpiTile& tl, int ch,
                                                    UpRegs* up = nullptr, const float* next = nullptr, int next_groups = 0) {
    {
        constexpr bool kVy = kStore == 1;
        const int cbase = ch * 64;
        // kUp: lanes 2i and 2i+1 are the two x-neighbours of ONE half-resolution pixel (tiles start at even pixels, W is even), i.e.
        // they need the same 64 bytes per 16-channel group: each holds HALF of them (32 B) and the pair swaps halves with shuffles.
        const int up_half = (e.row & 1) * 2;                 // float4 index of this lane's half
        if (e.leader) tma_store_wait_read<0>();            // previous store of this group has left the staging buffer
        group_bar_sync(e.bar_id);
        if (e.has_res) {
            if (e.leader) {
                mbar_expect_tx(e.res_bar, e.out_bytes);
                if (kVy && tl.y0 + 16 > tl.vh) {
                    // the tile spans two images: two rows at a time (a second full BOX would zero-fill what the first one loaded;
                    // y0 and vh are even, so a row pair stays in one image)
                    for (int j = 0; j < 16; j += 2) {
                        const int y = tl.y0 + j, wrap = y >= tl.vh ? 1 : 0;
                        tma_load_4d(e.s_out + (uint32_t)j * 1024u, e.tm_res_row, e.res_bar, tl.n0 + cbase, tl.x0, y - wrap * tl.vh, tl.img + wrap);
                    }
                } else {
                    tma_load_4d(e.s_out, e.tm_res, e.res_bar, tl.n0 + cbase, tl.x0, tl.y0, tl.img);
                }
            }
            mbar_wait(e.res_bar, res_phase);
            res_phase ^= 1u;
        }
        const uint32_t line = e.s_out + (uint32_t)e.row * 128u;
        // Four groups of 16 columns, software-pipelined: the TMEM load of group q+1 is in flight while group q is evaluated, and a
        // group's bias values are fetched BEFORE its wait (the volatile tcgen05 statements pin what follows them: a bias load
        // after the wait exposes its shared-memory latency once per group).  bias / SiLU run as packed fp32 pairs (fma.rn.f32x2:
        // the same two IEEE fmas as the scalar form, half the issue slots).
        // (the up-add variant keeps one buffer: a second one spills and costs 25 % of the layer)
        uint32_t va[16], vb[kUp ? 1 : 16];
        if (!kUp) tmem_ld16(t_row + (uint32_t)cbase, va);
        #pragma unroll
        for (int q16 = 0; q16 < 4; ++q16) {
            const int c0 = cbase + q16 * 16;
            if (c0 >= e.block_n) break;
            float4 b4[4];
            {
                const float4* bq = reinterpret_cast<const float4*>(e.s_bias + tl.n0 + c0);
                #pragma unroll
                for (int j = 0; j < 4; ++j) b4[j] = bq[j];               // act: 0.5 * bias
            }
            if (kUp) tmem_ld16(t_row + (uint32_t)c0, va);
            tmem_ld_wait();
            uint32_t (&v)[16] = (!kUp && (q16 & 1)) ? *reinterpret_cast<uint32_t (*)[16]>(&vb[0]) : va;
            if constexpr (!kUp) {
                if (q16 < 3 && c0 + 16 < e.block_n) tmem_ld16(t_row + (uint32_t)(c0 + 16), (q16 & 1) ? va : *reinterpret_cast<uint32_t (*)[16]>(&vb[0]));
            }
            if (kUp) {
                #pragma unroll
                for (int j = 0; j < 2; ++j) {
                    float4 o;                                // the partner's half
                    const float4 mine = up->u[2 * q16 + j];
                    o.x = __shfl_xor_sync(0xffffffffu, mine.x, 1); o.y = __shfl_xor_sync(0xffffffffu, mine.y, 1);
                    o.z = __shfl_xor_sync(0xffffffffu, mine.z, 1); o.w = __shfl_xor_sync(0xffffffffu, mine.w, 1);
                    const bool hi_mine = up_half != 0;       // selects keep v[] statically indexed (registers)
                    const float4 lo = hi_mine ? o : mine, hi = hi_mine ? mine : o;
                    v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + lo.x);
                    v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + lo.y);
                    v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + lo.z);
                    v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + lo.w);
                    v[4 * j + 8] = __float_as_uint(__uint_as_float(v[4 * j + 8]) + hi.x);
                    v[4 * j + 9] = __float_as_uint(__uint_as_float(v[4 * j + 9]) + hi.y);
                    v[4 * j + 10] = __float_as_uint(__uint_as_float(v[4 * j + 10]) + hi.z);
                    v[4 * j + 11] = __float_as_uint(__uint_as_float(v[4 * j + 11]) + hi.w);
                    if (q16 < next_groups) up->u[2 * q16 + j] = __ldg(reinterpret_cast<const float4*>(next) + 4 * q16 + up_half + j);
                }
            }
            float f[16];
            if (e.act) {
                #pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 hb = b4[j >> 1];
                    const unsigned long long h = epi_fma2(epi_pack(v[2 * j], v[2 * j + 1]), 0x3f0000003f000000ull,
                                                          (j & 1) ? epi_pack(__float_as_uint(hb.z), __float_as_uint(hb.w))
                                                                  : epi_pack(__float_as_uint(hb.x), __float_as_uint(hb.y)));
                    uint32_t h0, h1;
                    epi_unpack(h, h0, h1);
                    float t0, t1;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(__uint_as_float(h0)));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(__uint_as_float(h1)));
                    uint32_t r0, r1;
                    epi_unpack(epi_fma2(h, epi_pack(__float_as_uint(t0), __float_as_uint(t1)), h), r0, r1);
                    f[2 * j] = __uint_as_float(r0); f[2 * j + 1] = __uint_as_float(r1);
                }
            } else {
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4[j].x;
                    f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4[j].y;
                    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4[j].z;
                    f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4[j].w;
                }
            }
            #pragma unroll
            for (int q = 0; q < 2; ++q) {                    // 16 columns = 2 x 16 B chunks of the swizzled 128 B line
                const uint32_t addr = line + (((uint32_t)(q16 * 2 + q) ^ (uint32_t)(e.row & 7)) << 4);
                if (e.has_res) {
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
        if (kUp) {                                           // a partial chunk (fewer than four groups) refilled only its own groups
            #pragma unroll
            for (int q16 = 0; q16 < 4; ++q16)
                if (q16 < next_groups && cbase + q16 * 16 >= e.block_n) {
                    up->u[2 * q16] = __ldg(reinterpret_cast<const float4*>(next) + 4 * q16 + up_half);
                    up->u[2 * q16 + 1] = __ldg(reinterpret_cast<const float4*>(next) + 4 * q16 + up_half + 1);
                }
        }
        fence_proxy_async_smem();                            // generic-proxy writes -> async proxy (TMA)
        group_bar_sync(e.bar_id);
        if (e.leader) {
            tma_store_4d(e.tm_y, e.s_out, tl.n0 + cbase, tl.x0, tl.y0, tl.img);          // rows past the image (dummy rows, the next image's) are clipped
            if (kStore == 2) tma_store_4d(e.tm_y, e.s_out + 64u * 128u, tl.n0 + cbase, tl.x1, tl.y1, tl.img1);
            if (kVy)                                         // the next image's rows, two by two (TMA stores reject negative coordinates)
                for (int j = tl.vh - tl.y0; j < 16; j += 2)
                    tma_store_4d(e.tm_y_row, e.s_out + (uint32_t)j * 1024u, tl.n0 + cbase, tl.x0, j - (tl.vh - tl.y0), tl.img + 1);
            tma_store_commit();
        }
    }
}

// fp32 output (the head's raw logits) of ONE 32-channel chunk: same staging scheme, the tensor map is
// FLOAT32 with 32-element (128-byte) rows.
__device__ __forceinline__ void epilogue_chunk_f32(const EpiShared& e, uint32_t t_row, const EpiTile& tl, int ch) {
    const int cbase = ch * 32;
    if (e.leader) tma_store_wait_read<0>();
    group_bar_sync(e.bar_id);
    const uint32_t line = e.s_out + (uint32_t)e.row * 128u;
    #pragma unroll 1
    for (int q16 = 0; q16 < 2; ++q16) {
        const int c0 = cbase + q16 * 16;
        if (c0 >= e.block_n) break;
        uint32_t v[16];
        tmem_ld16(t_row + (uint32_t)c0, v);
        tmem_ld_wait();
        const float4* bq = reinterpret_cast<const float4*>(e.s_bias + tl.n0 + c0);
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 b4 = bq[j];
            float f0 = __uint_as_float(v[4 * j + 0]), f1 = __uint_as_float(v[4 * j + 1]);
            float f2 = __uint_as_float(v[4 * j + 2]), f3 = __uint_as_float(v[4 * j + 3]);
            if (e.act) {
                f0 = silu_from_half(fmaf(f0, 0.5f, b4.x)); f1 = silu_from_half(fmaf(f1, 0.5f, b4.y));
                f2 = silu_from_half(fmaf(f2, 0.5f, b4.z)); f3 = silu_from_half(fmaf(f3, 0.5f, b4.w));
            } else {
                f0 += b4.x; f1 += b4.y; f2 += b4.z; f3 += b4.w;
            }
            const uint32_t addr = line + (((uint32_t)(q16 * 4 + j) ^ (uint32_t)(e.row & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(f0), "f"(f1), "f"(f2), "f"(f3) : "memory");
        }
    }
    fence_proxy_async_smem();
    group_bar_sync(e.bar_id);
    if (e.leader) {
        tma_store_4d(e.tm_y, e.s_out, tl.n0 + cbase, tl.x0, tl.y0, tl.img);
        tma_store_commit();
    }
}

}  // namespace tc
}  // namespace yms
