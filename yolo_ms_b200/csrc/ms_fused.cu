// MS-Block branch layer on one SM pass: depthwise k x k (CUDA cores, packed fp32 FMAs) fused with the 1x1
// convolutions around it (tcgen05), so that the EXPANDED tensor of the layer never travels to HBM.
//
// The repo-local MS-Block layer (yolo_ms_b200/modules.py::MSBlock, built from the reference's Conv unit
// yolov8/model/components.py:69-77 with its `groups` argument) is
//        x [c] --pw1 1x1--> e [E = 2c] --depthwise k x k--> d [E] --pw2 1x1--> y [c]
// each step followed by the folded BN bias and SiLU.  Unfused it moves 11c elements per pixel through HBM
// (the two expanded tensors are written once and read once each); the kernel below has three modes:
//   mode 0  depthwise only            e (TMA halo tiles) -> d -> TMA store            (yms_dwconv: Conv(c, c, k, groups=c))
//   mode 1  depthwise -> pw2          e (TMA halo tiles) -> d (smem, UMMA A operand) -> tcgen05.mma -> epilogue -> y
//   mode 2  pw1 -> depthwise -> pw2   x halo tile (TMA) -> tcgen05.mma -> e (TMEM -> bias/SiLU -> smem, zero outside the image)
//                                     -> d -> tcgen05.mma -> epilogue -> y        (3c elements per pixel)
// Work decomposition: output tile = 16 x 8 (or 24 x 5) pixels of one image (M = 128 rows of the pw2 GEMM); the expanded channels
// are walked in chunks of 64 (one 128-byte swizzle row); chunk n of pw2's K dimension is produced by the depthwise stage while the
// tensor core consumes chunk n-1.  Mode 0 has no GEMM: its work unit is ONE (tile, chunk) pair, so that a 20 x 20 map with 512
// channels spreads over all SMs.
//
// Roles (704 threads, one CTA per SM): warp 0 TMA producer | warp 1 MMA issuer (mode 0: issues the TMA stores) | warps 2-5 pw2
// epilogue | warps 6-21 compute warps.  The depthwise stage is the CUDA-core bound of the layer, and it is latency-bound, not
// throughput-bound, when whole warp groups march through a chunk in lock step (round 2, first version: two groups of 8 warps with
// named barriers; ncu: the depthwise instructions were 28 % of the groups' samples, the rest waits).  So the unit of work is a
// WARP ITEM -- a pixel block of one chunk, lane = channel pair -- and items of consecutive chunks are dealt round-robin to the
// depthwise warps, each of which runs its own sequence against the chunk mbarriers (e_full / d_empty in, d_full / e_empty out; no
// named barrier anywhere).  In mode 2 compute warps 0-7 are the pw1 epilogue (two per TMEM lane quadrant) and run one e stage
// ahead of the 8 depthwise warps.
//
// Depthwise items (fp32 accumulation throughout -- the repo's numeric contract; FFMA2 = fma.rn.f32x2 on a channel pair):
//   * halo rows live in shared memory at a pitch of 24 / 32 pixels (a multiple of 8), 128 B per pixel, 16-byte chunks XOR-swizzled
//     by (pixel & 7) -- exactly what TMA SWIZZLE_128B writes when every halo row is its own box at a 3072-byte pitch.  A warp
//     reads ONE pixel per LDS.32 (its 32 lanes are the 32 channel pairs = the pixel's 128 contiguous bytes: conflict-free).
//   * k = 3, 16 x 8 tiles: item = 4 x 4 output pixels; the 6 x 6 halo is loaded and converted once, the 9 taps sit in registers.
//   * other k / tiles: item = 2 rows x 8 pixels; every input row is loaded and converted once and feeds both output rows
//     (2k FFMA2 per loaded pixel); the weights of two kernel rows sit in registers (a ring over the input rows).  Shared-memory
//     traffic per chunk drops from k x to (k+1)/2 x the halo tile against one-row strips -- the strips were bandwidth-bound.
//   * the result (bias, SiLU, bf16) is written as a [128 x 64] K-major SWIZZLE_128B tile = the A operand of pw2's MMAs
//     (or the source of a TMA store in mode 0).
#include "conv_plan.h"

#include <string.h>
#include <new>

namespace yms {
namespace {
using namespace tc;

// Tile geometry: output pixels per tile (TW x TH <= 128 = the M of the pw2 GEMM), and the pitch (pixels, a multiple of 8)
// at which halo rows sit in shared memory.  GEOM 0 = 16 x 8 (pitch 24); GEOM 1 = 24 x 5 (pitch 32) for maps whose width is
// a bad fit for 16-pixel tiles (a 20 x 20 map is 6 tiles at 52 % with GEOM 0, 4 tiles at 83 % with GEOM 1).
template <int GEOM> struct Geom;
template <> struct Geom<0> { static constexpr int TW = 16, TH = 8; };
template <> struct Geom<1> { static constexpr int TW = 24, TH = 5; };
// Halo-row pitch in pixels: TW + k - 1 rounded up to a multiple of 8, except k = 5 (20 / 28 pixels: a multiple of 4 -- the swizzle
// phase of a pixel is then (row * pitch + x) & 7, still a compile-time constant per tap for items that start on even rows; the
// 17 % smaller stage is what lets the dw -> pw2 mode of the 256-channel layers keep three stages beside its resident weights)
__host__ __device__ constexpr int halo_pitch(int k, int geom) {
    return k == 5 ? (geom ? 28 : 20) : (((geom ? 24 : 16) + k - 1 + 7) & ~7);
}

constexpr int kMaxE = 4;                         // e-ring stages
constexpr int kDTile = 128 * 128;                // one [128 x 64] bf16 tile
constexpr int kFirstCw = 6;                      // warp 0 producer, 1 MMA / storer, 2-5 pw2 epilogue, 6.. compute warps
constexpr int kCwCount = 16;                     // compute warps
constexpr int kPWarps = 8;                       // mode 2: compute warps 0..7 = pw1 epilogue, the rest depthwise
constexpr int kG1Warp = kFirstCw + kCwCount;     // mode 2: issuer of the pw1 MMAs (its own warp: never blocked behind the pw2 chain)
constexpr int kThreads = (kG1Warp + 1) * 32;
constexpr int kMaxD = 4;                         // d-ring buffers
constexpr int kSmemLimit = 232448;

// Depthwise warp items of one chunk (host and device must agree: the chunk barriers count item arrivals)
__host__ __device__ constexpr bool item_is_block(int k, int geom) { return k == 3 && geom == 0; }
__host__ __device__ constexpr int item_rows(int k) { return k <= 7 ? 2 : 1; }
__host__ __device__ constexpr int items_per_chunk(int k, int geom) {
    return item_is_block(k, geom) ? 8
         : (geom ? (24 / 8) * ((5 + item_rows(k) - 1) / item_rows(k)) : (16 / 8) * ((8 + item_rows(k) - 1) / item_rows(k)));
}

struct MsParams {
    int mode, ksize, geom;
    int tiles_x, tiles_y, batch, total_tiles;
    int units, cpu;                              // work units of the grid and chunks per unit (mode 0: tiles * chunks, 1; else tiles, chunks)
    uint32_t mg_tiles_x, mg_tiles_y;
    int e_ch, n_chunks, tail_ksteps;             // expanded channels, 64-channel chunks, k-steps of the last chunk
    int c_out, block_n, act2;
    int e_stages, n_items, d_bufs, g1_stages;
    uint32_t p_mask; int n_p;                    // mode 2: compute warps that run the pw1 epilogue (the others are depthwise warps)
    int w2_resident, w2_tile_bytes;
    int acc_stride, tmem_cols;
    int stage_bytes, off_dww, off_dwb, off_w2;   // layout of one e stage
    int so_x, so_e, so_d, so_w2, so_out, so_w1, so_bias;   // shared-memory carve-up (bytes from the 1024-aligned base)
    int x_kb_stride;                             // mode 2: bytes between the halo tiles of consecutive K blocks (hp * 128 rounded up to 1 KB)
    int x_bufs, x_buf_bytes;                     // mode 2: 1 or 2 buffers for the input halo tile
    uint32_t stage_tx;                           // bytes TMA delivers per stage
    int bias_pad;
    const float* bias2;
    // mode 2
    int c_in1, c_in2, kb1, kb2;                  // pw1 sources (K-concatenated), 64-channel blocks of each
    int mt;                                      // M tiles (128 halo pixels each) of the pw1 GEMM
    int hp;                                      // halo pixels = (TW+k-1) * (TH+k-1)
    int img_w, img_h;
    int w1_tile_bytes;                           // one [64 x 64] weight tile = 8192
    int g1_stride;                               // TMEM columns between the two pw1 accumulator stages
    int g1_base;                                 // first TMEM column of the pw1 accumulators
    const float* bias1;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ unsigned long long pack64(uint32_t lo, uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack64(unsigned long long v, float& lo, float& hi) {
    uint32_t a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
    lo = __uint_as_float(a); hi = __uint_as_float(b);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// bf16x2 word -> packed (f32, f32): the low half through PRMT (ALU pipe; a shift would be emitted as an IMAD on the FMA pipe,
// which is the pipe the depthwise stage is bound by), the high half through a mask
__device__ __forceinline__ unsigned long long bf16x2_to_f32x2(uint32_t u) {
    uint32_t lo;
    asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(u));
    return pack64(lo, u & 0xffff0000u);
}

struct TileXY { int img, x0, y0; };
template <int GEOM>
__device__ __forceinline__ TileXY decode_tile(const MsParams& p, int t) {
    TileXY c;
    uint32_t q = fast_div((uint32_t)t, p.mg_tiles_x);
    const int tx = t - (int)q * p.tiles_x;
    const uint32_t q2 = fast_div(q, p.mg_tiles_y);
    const int ty = (int)(q - q2 * p.tiles_y);
    c.img = (int)q2; c.x0 = tx * Geom<GEOM>::TW; c.y0 = ty * Geom<GEOM>::TH;
    return c;
}

// SiLU of two packed fp32 values given as acc (pre-activation without bias) and hb = 0.5 * bias:  h = 0.5 * acc + hb,
// silu = h + h * tanh(h)  (conv_epilogue.cuh's one-MUFU form), both FMAs packed -> bf16x2
__device__ __forceinline__ uint32_t silu2_bf16(unsigned long long acc, unsigned long long half2, unsigned long long hb) {
    const unsigned long long h = fma2(acc, half2, hb);
    float h0, h1, t0, t1;
    unpack64(h, h0, h1);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
    float r0, r1;
    unpack64(fma2(h, pack64(__float_as_uint(t0), __float_as_uint(t1)), h), r0, r1);
    return pack_bf16x2(r0, r1);
}
constexpr unsigned long long kHalf2 = 0x3f0000003f000000ull;     // (0.5f, 0.5f)

// k = 3, 16 x 8 tiles: warp item = a 4 x 4 block of output pixels, lane = channel pair.  Every input pixel of the block's 6 x 6
// halo is loaded (LDS.32: the 32 lanes of a warp are the 32 channel pairs of ONE pixel = its 128 contiguous bytes) and converted
// ONCE and feeds up to 9 FFMA2; the 9 taps sit in registers.
//   e_blk = stage + ((by*4) * PITCH + bx*4) * 128, cp = channel pair (lane), xs = (bx*4) & 7 (0 or 4)
template <int PITCH, int TW>
__device__ __forceinline__ void dw_block_k3(uint32_t e_blk, uint32_t w_base, uint32_t b_addr, uint32_t d_blk, int cp, int xs) {
    const uint32_t sub = (uint32_t)((cp & 3) * 4), chunk = (uint32_t)(cp >> 2);
    uint32_t eb[6];
    #pragma unroll
    for (int j = 0; j < 6; ++j) eb[j] = e_blk + (uint32_t)(j * 128) + ((chunk ^ (uint32_t)((xs + j) & 7)) << 4) + sub;
    unsigned long long w[9];
    #pragma unroll
    for (int t = 0; t < 9; ++t) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(w[t]) : "r"(w_base + (uint32_t)(t * 256)));
    unsigned long long acc[4][4];
    #pragma unroll
    for (int iy = 0; iy < 6; ++iy) {
        unsigned long long v[6];
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
            uint32_t u;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(eb[j] + (uint32_t)(iy * PITCH * 128)));
            v[j] = bf16x2_to_f32x2(u);
        }
        #pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int oy = iy - ky;
            if (oy >= 0 && oy < 4) {
                #pragma unroll
                for (int ox = 0; ox < 4; ++ox) {
                    #pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        acc[oy][ox] = (ky == 0 && kx == 0) ? fma2(v[ox], w[0], 0ull) : fma2(v[ox + kx], w[ky * 3 + kx], acc[oy][ox]);
                }
            }
        }
    }
    unsigned long long hb;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(hb) : "r"(b_addr));
    hb = fma2(hb, kHalf2, 0ull);
    #pragma unroll
    for (int ox = 0; ox < 4; ++ox) {
        const uint32_t db = d_blk + (uint32_t)(ox * 128) + ((chunk ^ (uint32_t)((xs + ox) & 7)) << 4) + sub;
        #pragma unroll
        for (int oy = 0; oy < 4; ++oy) {
            const uint32_t o = silu2_bf16(acc[oy][ox], kHalf2, hb);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(db + (uint32_t)(oy * TW * 128)), "r"(o) : "memory");
        }
    }
}

// General warp item: NR output rows x 8 pixels (x0 a multiple of 8), lane = channel pair.  Input row iy of the item's halo
// (NR + K - 1 rows of 8 + K - 1 pixels) is loaded and converted once; it is kernel row ky = iy - r of output row r, so the weights
// of kernel rows iy and iy - 1 are live (a two-slot ring, K LDS.64 per input row).
//   e_item = stage + (oy0 * PITCH + x0) * 128 with (oy0 * PITCH + x0) % 8 == 0;  d_item = d tile + (oy0 * TW + x0) * 128;
//   w_lane / b_lane = this lane's 8 bytes of tap 0 / of the bias.
template <int K, int PITCH, int TW, int NR>
__device__ __forceinline__ void dw_rows(uint32_t e_item, uint32_t w_lane, uint32_t b_lane, uint32_t d_item, int cp) {
    static_assert(NR == 1 || NR == 2, "row ring holds two kernel rows");
    const uint32_t sub = (uint32_t)((cp & 3) * 4), chunk = (uint32_t)(cp >> 2);
    uint32_t sw[8];                                   // lane part of the swizzled address of a pixel with (pixel & 7) == i
    #pragma unroll
    for (int i = 0; i < 8; ++i) sw[i] = ((chunk ^ (uint32_t)i) << 4) + sub;
    unsigned long long acc[NR][8];
    unsigned long long w[2][K];
    #pragma unroll
    for (int iy = 0; iy < NR + K - 1; ++iy) {
        if (iy < K) {
            #pragma unroll
            for (int kx = 0; kx < K; ++kx)
                asm volatile("ld.shared.u64 %0, [%1];" : "=l"(w[iy & 1][kx]) : "r"(w_lane + (uint32_t)((iy * K + kx) * 256)));
        }
        #pragma unroll
        for (int j = 0; j < 8 + K - 1; ++j) {
            uint32_t u;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u) : "r"(e_item + sw[(iy * PITCH + j) & 7] + (uint32_t)((iy * PITCH + j) * 128)));
            const unsigned long long v = bf16x2_to_f32x2(u);
            #pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int ky = iy - r;
                if (ky >= 0 && ky < K) {
                    #pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        const int o = j - kx;
                        if (o >= 0 && o < 8)
                            acc[r][o] = (ky == 0 && kx == 0) ? fma2(v, w[0][0], 0ull) : fma2(v, w[ky & 1][kx], acc[r][o]);
                    }
                }
            }
        }
    }
    unsigned long long hb;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(hb) : "r"(b_lane));
    hb = fma2(hb, kHalf2, 0ull);
    #pragma unroll
    for (int r = 0; r < NR; ++r) {
        #pragma unroll
        for (int o = 0; o < 8; ++o) {
            const uint32_t v = silu2_bf16(acc[r][o], kHalf2, hb);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(d_item + sw[o] + (uint32_t)((r * TW + o) * 128)), "r"(v) : "memory");
        }
    }
}

template <int K, int GEOM>
__global__ void __launch_bounds__(kThreads, 1)
ms_layer_kernel(const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_dww,
                const __grid_constant__ CUtensorMap tm_dwb, const __grid_constant__ CUtensorMap tm_w2,
                const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_x,
                const __grid_constant__ CUtensorMap tm_x2, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ MsParams p) {
    constexpr int TW = Geom<GEOM>::TW, TH = Geom<GEOM>::TH, PITCH = halo_pitch(K, GEOM);
    constexpr int HWX = TW + K - 1, HWY = TH + K - 1, PAD = K / 2;
    constexpr bool kBlock = item_is_block(K, GEOM);
    constexpr int NR = item_rows(K);
    constexpr int kStripsX = TW / 8;
    constexpr int kItems = items_per_chunk(K, GEOM);
    static_assert(HWX <= PITCH && TW * TH <= 128 && TW % 8 == 0 && (PITCH % 8 == 0 || (PITCH % 4 == 0 && !kBlock)), "tile geometry");
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gbase = smem_dyn + (base - smem_u32(smem_dyn));
    // carve-up (host: yms_ms_plan_create): [mode 2: x halo tiles][e stages][d x 2][W2 resident][staging][mode 2: W1][bias][barriers].
    // The x tiles come first: the last M tile of a K block reads 128 rows even when the halo has fewer, i.e. past the tile
    // into whatever follows (harmless garbage rows of the accumulator) -- that must still be inside the allocation.
    const uint32_t s_x = base + (uint32_t)p.so_x;
    const uint32_t s_e0 = base + (uint32_t)p.so_e;
    const uint32_t s_d0 = base + (uint32_t)p.so_d;
    const uint32_t s_w2 = base + (uint32_t)p.so_w2;
    const uint32_t s_out = base + (uint32_t)p.so_out;
    const uint32_t s_w1 = base + (uint32_t)p.so_w1;
    const uint32_t w1_bytes = p.mode == 2 ? (uint32_t)(p.n_chunks * (p.kb1 + p.kb2) * p.w1_tile_bytes) : 0u;
    const uint32_t s_biasu = base + (uint32_t)p.so_bias;
    float* s_bias = reinterpret_cast<float*>(gbase + (s_biasu - base));          // pw2 bias (0.5 x when act), then pw1 bias (mode 2)
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_bias) + p.bias_pad * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto e_full = [&](int s) { return bar0 + 8u * s; };                          // chunk inputs of stage s complete (TMA + pw1 epilogue warps)
    auto e_empty = [&](int s) { return bar0 + 8u * (kMaxE + s); };               // every depthwise item of the stage's chunk has been computed
    auto d_full = [&](int s) { return bar0 + 8u * (8 + s); };                    // every item of the chunk has been written to d buffer s
    auto d_empty = [&](int s) { return bar0 + 8u * (12 + s); };                  // the MMAs (mode 0: the TMA store) that read d buffer s are done
    auto acc_full = [&](int s) { return bar0 + 8u * (16 + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (18 + s); };
    const uint32_t res_bar = bar0 + 8u * 20;
    const uint32_t w_bar = bar0 + 8u * 21;
    auto g1_full = [&](int s) { return bar0 + 8u * (22 + s); };                  // mode 2: pw1 accumulator stage ready (up to 3 stages)
    auto g1_empty = [&](int s) { return bar0 + 8u * (25 + s); };
    auto x_full = [&](int s) { return bar0 + 8u * (28 + s); };                   // mode 2: x halo tile landed / may be overwritten
    auto x_empty = [&](int s) { return bar0 + 8u * (30 + s); };
    constexpr int kNumBars = 32;
    static_assert(kMaxE == 4 && kMaxD == 4, "barrier map");
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tm_dww); prefetch_tmap(&tm_dwb); prefetch_tmap(&tm_y);
            if (p.mode != 2) prefetch_tmap(&tm_e);
            if (p.mode >= 1) prefetch_tmap(&tm_w2);
            if (p.mode == 2) { prefetch_tmap(&tm_x); prefetch_tmap(&tm_w1); if (p.kb2) prefetch_tmap(&tm_x2); }
        }
        if (lane < kNumBars) {
            uint32_t cnt = 1;
            if (lane < kMaxE) cnt = 1u + (uint32_t)p.n_p;                              // e_full: producer (+ pw1 epilogue warps)
            else if (lane < 2 * kMaxE) cnt = (uint32_t)kItems + ((p.mode >= 1 && !p.w2_resident) ? 1u : 0u);  // e_empty: items (+ MMA commit)
            else if (lane < 12) cnt = (uint32_t)kItems;                                                       // d_full: items
            else if (lane >= 18 && lane < 20) cnt = 4u;                                                       // acc_empty: 4 epilogue warps
            else if (lane >= 25 && lane < 28) cnt = (uint32_t)max(p.n_p, 1);                                        // g1_empty: pw1 epilogue warps
            mbar_init(bar0 + 8u * lane, cnt);
        }
        fence_barrier_init();
        __syncwarp();
        if (p.mode >= 1 && elect_one()) {                      // constants of the program: fetched before the grid dependency resolves
            uint32_t tx = 0;
            if (p.w2_resident) tx += (uint32_t)(p.n_chunks * p.w2_tile_bytes);
            if (p.mode == 2) tx += w1_bytes;
            if (tx) {
                mbar_expect_tx(w_bar, tx);
                if (p.w2_resident)
                    for (int j = 0; j < p.n_chunks; ++j) tma_load_3d(s_w2 + (uint32_t)(j * p.w2_tile_bytes), &tm_w2, w_bar, j * 64, 0, 0);
                if (p.mode == 2) {
                    const int kbt = p.kb1 + p.kb2;
                    for (int j = 0; j < p.n_chunks; ++j)
                        for (int kb = 0; kb < kbt; ++kb)
                            tma_load_3d(s_w1 + (uint32_t)((j * kbt + kb) * p.w1_tile_bytes), &tm_w1, w_bar,
                                        kb < p.kb1 ? kb * 64 : p.c_in1 + (kb - p.kb1) * 64, j * 64, 0);
                }
            }
        }
        __syncwarp();
    }
    if (warp == 1 && p.mode >= 1) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
    for (int i = threadIdx.x; i < p.bias_pad; i += blockDim.x) {
        float v = 0.f;
        if (p.mode >= 1 && i < p.c_out) v = p.act2 ? 0.5f * p.bias2[i] : p.bias2[i];
        else if (p.mode == 2 && i >= 256 && i - 256 < p.e_ch) v = 0.5f * p.bias1[i - 256];
        s_bias[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = p.mode >= 1 ? *tmem_slot : 0u;
    pdl_launch_dependents();
    pdl_wait();

    const int grid = (int)gridDim.x;
    // unit u of the grid: mode 0 = chunk (u % n_chunks) of tile (u / n_chunks); else all chunks of tile u
    auto unit_tile = [&](int u, int& j0) { int t = u; j0 = 0; if (p.mode == 0) { t = u / p.n_chunks; j0 = u - t * p.n_chunks; } return t; };

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int s = 0; uint32_t ph = 0;                        // e-ring stage / phase of the next chunk
            int xb = 0; uint32_t xph = 0;                      // x buffer / phase of the next input halo tile
            const int kbt = p.kb1 + p.kb2;
            auto load_x = [&](int t) {
                const TileXY tc = decode_tile<GEOM>(p, t);
                mbar_wait_sleep(x_empty(xb), xph ^ 1u);
                mbar_expect_tx(x_full(xb), (uint32_t)(kbt * HWX * HWY * 128));
                for (int kb = 0; kb < kbt; ++kb)
                    tma_load_4d(s_x + (uint32_t)(xb * p.x_buf_bytes + kb * p.x_kb_stride), kb < p.kb1 ? &tm_x : &tm_x2, x_full(xb),
                                (kb < p.kb1 ? kb : kb - p.kb1) * 64, tc.x0 - PAD, tc.y0 - PAD, tc.img);
                if (++xb == p.x_bufs) { xb = 0; xph ^= 1u; }
            };
            if (p.mode == 2 && (int)blockIdx.x < p.units) load_x(blockIdx.x);
            for (int u = blockIdx.x; u < p.units; u += grid) {
                int j0;
                const int t = unit_tile(u, j0);
                const TileXY tc = decode_tile<GEOM>(p, t);
                // the NEXT tile's input halo goes first: with two buffers it lands while this tile is processed; with one
                // buffer it is issued after this tile's chunk loads (it has to wait for this tile's last pw1 MMA anyway)
                if (p.mode == 2 && p.x_bufs == 2 && u + grid < p.units) load_x(u + grid);
                for (int jj = 0; jj < p.cpu; ++jj) {
                    const int j = j0 + jj;
                    mbar_wait_sleep(e_empty(s), ph ^ 1u);
                    const uint32_t st = s_e0 + (uint32_t)(s * p.stage_bytes);
                    mbar_expect_tx(e_full(s), p.stage_tx);
                    if (p.mode != 2) {
                        if constexpr (PITCH == HWX) {          // rows are contiguous: the whole halo is one box (host: box height HWY)
                            tma_load_4d(st, &tm_e, e_full(s), j * 64, tc.x0 - PAD, tc.y0 - PAD, tc.img);
                        } else {
                            #pragma unroll 1
                            for (int hy = 0; hy < HWY; ++hy)
                                tma_load_4d(st + (uint32_t)(hy * PITCH * 128), &tm_e, e_full(s), j * 64, tc.x0 - PAD, tc.y0 - PAD + hy, tc.img);
                        }
                    }
                    tma_load_2d(st + (uint32_t)p.off_dww, &tm_dww, e_full(s), j * 64, 0);
                    tma_load_2d(st + (uint32_t)p.off_dwb, &tm_dwb, e_full(s), j * 64, 0);
                    if (p.mode >= 1 && !p.w2_resident) tma_load_3d(st + (uint32_t)p.off_w2, &tm_w2, e_full(s), j * 64, 0, 0);
                    if (++s == p.e_stages) { s = 0; ph ^= 1u; }
                }
                if (p.mode == 2 && p.x_bufs == 1 && u + grid < p.units) load_x(u + grid);
            }
        }
    } else if (warp == 1) {
        if (p.mode == 0) {
            // ================= mode 0: TMA store of every finished d tile (one elected thread) =================
            if (elect_one()) {
                int ds = 0; uint32_t dph = 0;
                for (int u = blockIdx.x; u < p.units; u += grid) {
                    int j0;
                    const int t = unit_tile(u, j0);
                    const TileXY tc = decode_tile<GEOM>(p, t);
                    mbar_wait_sleep(d_full(ds), dph);                        // the writers fenced (generic -> async proxy) before arriving
                    tma_store_4d(&tm_y, s_d0 + (uint32_t)ds * kDTile, j0 * 64, tc.x0, tc.y0, tc.img);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(d_empty(ds));
                    if (++ds == p.d_bufs) { ds = 0; dph ^= 1u; }
                }
            }
        } else if (elect_one()) {
            // ================= MMA issuer (one elected thread) =================
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            if (p.w2_resident) mbar_wait_sleep(w_bar, 0u);          // armed in the prologue
            tc_fence_after();
            uint32_t ti = 0;
            int s = 0; uint32_t ph = 0;                        // e-ring stage / phase of the next chunk
            int ds = 0; uint32_t dph = 0;                      // d-ring buffer / phase of the next chunk
            for (int t = blockIdx.x; t < p.total_tiles; t += grid, ++ti) {
                const int as = (int)(ti & 1u);
                mbar_wait_sleep(acc_empty(as), ((ti >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.acc_stride);
                for (int j = 0; j < p.n_chunks; ++j) {
                    if (!p.w2_resident) mbar_wait_sleep(e_full(s), ph);
                    mbar_wait_sleep(d_full(ds), dph);
                    tc_fence_after();
                    const uint32_t a16 = ((s_d0 + (uint32_t)ds * kDTile) & 0x3FFFFu) >> 4;
                    const uint32_t b16 = ((p.w2_resident ? s_w2 + (uint32_t)(j * p.w2_tile_bytes)
                                                         : s_e0 + (uint32_t)(s * p.stage_bytes + p.off_w2)) & 0x3FFFFu) >> 4;
                    const int ks = (j == p.n_chunks - 1) ? p.tail_ksteps : 4;
                    for (int k = 0; k < ks; ++k)
                        umma_bf16(d_tmem, hi | (uint64_t)(a16 + 2 * k), hi | (uint64_t)(b16 + 2 * k), idesc2, (j | k) ? 1u : 0u);
                    umma_commit(d_empty(ds));
                    if (!p.w2_resident) umma_commit(e_empty(s));
                    if (j == p.n_chunks - 1) umma_commit(acc_full(as));
                    if (++s == p.e_stages) { s = 0; ph ^= 1u; }
                    if (++ds == p.d_bufs) { ds = 0; dph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp < kFirstCw) {
        // ================= pw2 epilogue: TMEM -> bias / SiLU -> bf16 -> swizzled staging -> TMA store =================
        if (p.mode >= 1) {
            EpiShared e;
            e.tm_y = &tm_y; e.tm_res = &tm_y;
            e.res_bar = res_bar;
            e.s_out = s_out;
            e.s_bias = s_bias;
            e.block_n = p.block_n; e.c_out = p.c_out; e.act = p.act2; e.has_res = 0;
            e.out_bytes = (uint32_t)kDTile;
            e.bar_id = 1;
            e.leader = (warp == 2) && lane == 0;
            e.row = (warp & 3) * 32 + lane;
            const int n_chunks_out = (p.block_n + 63) >> 6;
            uint32_t res_phase = 0u, ti = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += grid, ++ti) {
                const int as = (int)(ti & 1u);
                const TileXY tc = decode_tile<GEOM>(p, t);
                EpiTile tl; tl.n0 = 0; tl.x0 = tc.x0; tl.y0 = tc.y0; tl.img = tc.img;
                mbar_wait_sleep(acc_full(as), (ti >> 1) & 1u);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(as * p.acc_stride);
                for (int ch = 0; ch < n_chunks_out; ++ch) epilogue_chunk_bf16(e, res_phase, t_row, tl, ch);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(as));
            }
            if (e.leader) tma_store_wait_read<0>();
        }
    } else if (warp >= kFirstCw && warp < kFirstCw + kCwCount && ((p.p_mask >> (warp - kFirstCw)) & 1u)) {
        // ================= mode 2, pw1 epilogue warps: accumulator rows (halo pixels) -> bias / SiLU -> bf16 -> halo tile of the
        // chunk's e stage, zero outside the image.  Two warps per TMEM lane quadrant, dealt the M tiles of the halo by parity.
        const int wq = warp & 3, half = (warp - kFirstCw) >> 2;
        const int hp = p.hp, mt = p.mt, img_w = p.img_w, img_h = p.img_h, e_stages = p.e_stages, n_chunks = p.n_chunks;
        const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
        const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)p.g1_base;
        const uint32_t g1_stride = (uint32_t)p.g1_stride;
        const int g1_stages = p.g1_stages;
        int s = 0; uint32_t ph = 0;
        int gs = 0; uint32_t gph = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += grid) {
            const TileXY tc = decode_tile<GEOM>(p, t);
            for (int j = 0; j < n_chunks; ++j) {
                const uint32_t st = s_e0 + (uint32_t)s * stage_bytes;
                mbar_wait_sleep(e_empty(s), ph ^ 1u);                         // the depthwise items of the stage's previous chunk are done
                mbar_wait_sleep(g1_full(gs), gph);
                tc_fence_after();
                for (int m = half; m < mt; m += 2) {
                    if (m * 128 + wq * 32 >= hp) break;                       // the whole warp is past the halo
                    const int hpix = m * 128 + wq * 32 + lane;                // halo pixel of this thread = accumulator row
                    const int hy = hpix / HWX, hx = hpix - hy * HWX;
                    const int gx = tc.x0 - PAD + hx, gy = tc.y0 - PAD + hy;
                    const bool inside = hpix < hp && gx >= 0 && gx < img_w && gy >= 0 && gy < img_h;
                    const uint32_t t_row = t_lane + (uint32_t)gs * g1_stride + (uint32_t)(m * 64);
                    const uint32_t line = st + (uint32_t)((hy * PITCH + hx) * 128);
                    // 4 groups of 16 columns, software-pipelined: the TMEM load of group q+1 is in flight while group q is
                    // evaluated, and a group's 16 bias values are fetched BEFORE its wait (the volatile tcgen05 statements pin
                    // whatever follows them, so a bias load placed after the wait would expose its latency once per group)
                    const float4* bq = reinterpret_cast<const float4*>(s_bias + 256 + j * 64);
                    const uint32_t swz = (uint32_t)((hy * PITCH + hx) & 7);
                    const int nq = min(4, (p.e_ch - j * 64 + 15) >> 4);            // 16-column groups of real channels in this chunk
                    uint32_t va[16], vb[16];
                    tmem_ld16(t_row, va);
                    #pragma unroll
                    for (int q16 = 0; q16 < 4; ++q16) {
                        if (q16 >= nq) break;                                      // tail chunk (E % 64 != 0): the rest is zero padding of the weights
                        float4 hb[4];
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) hb[i] = bq[q16 * 4 + i];
                        tmem_ld_wait();
                        uint32_t (&v)[16] = (q16 & 1) ? vb : va;
                        if (q16 + 1 < nq) tmem_ld16(t_row + (uint32_t)((q16 + 1) * 16), (q16 & 1) ? va : vb);
                        if (inside) {
                            uint32_t o[8];
                            #pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                o[2 * i] = silu2_bf16(pack64(v[4 * i + 0], v[4 * i + 1]), kHalf2, pack64(__float_as_uint(hb[i].x), __float_as_uint(hb[i].y)));
                                o[2 * i + 1] = silu2_bf16(pack64(v[4 * i + 2], v[4 * i + 3]), kHalf2, pack64(__float_as_uint(hb[i].z), __float_as_uint(hb[i].w)));
                            }
                            #pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t addr = line + (((uint32_t)(q16 * 2 + h) ^ swz) << 4);
                                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o[4 * h]), "r"(o[4 * h + 1]), "r"(o[4 * h + 2]), "r"(o[4 * h + 3]) : "memory");
                            }
                        } else if (hpix < hp) {                               // halo pixels outside the image: the depthwise conv's zero padding
                            #pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t addr = line + (((uint32_t)(q16 * 2 + h) ^ swz) << 4);
                                asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(g1_empty(gs)); mbar_arrive(e_full(s)); }
                if (++s == e_stages) { s = 0; ph ^= 1u; }
                if (++gs == g1_stages) { gs = 0; gph ^= 1u; }
            }
        }
    } else if (warp == kG1Warp) {
        // ================= mode 2: issuer of the pw1 GEMMs (one elected thread).  It runs g1_stages chunks ahead of the pw1
        // epilogue, also across tile boundaries, and is throttled only by its accumulator stages and the x halo buffers (a buffer
        // is released by the last pw1 MMA of its tile) -- never by the depthwise / pw2 chain.
        if (p.mode == 2 && elect_one()) {
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            mbar_wait_sleep(w_bar, 0u);
            tc_fence_after();
            const int kbt = p.kb1 + p.kb2;
            const int tail1 = ((p.c_in1 - (p.kb1 - 1) * 64) + 15) >> 4;
            const int tail2 = p.kb2 ? (((p.c_in2 - (p.kb2 - 1) * 64) + 15) >> 4) : 4;
            int xb = 0; uint32_t xph = 0;
            int gs = 0; uint32_t gph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += grid) {
                mbar_wait_sleep(x_full(xb), xph);
                tc_fence_after();
                const uint32_t sx_cur = s_x + (uint32_t)(xb * p.x_buf_bytes);
                for (int j = 0; j < p.n_chunks; ++j) {
                    mbar_wait_sleep(g1_empty(gs), gph ^ 1u);
                    tc_fence_after();
                    for (int m = 0; m < p.mt; ++m) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(p.g1_base + gs * p.g1_stride + m * 64);
                        for (int kb = 0; kb < kbt; ++kb) {
                            const uint32_t a16 = ((sx_cur + (uint32_t)(kb * p.x_kb_stride + m * kDTile)) & 0x3FFFFu) >> 4;
                            const uint32_t b16 = ((s_w1 + (uint32_t)((j * kbt + kb) * p.w1_tile_bytes)) & 0x3FFFFu) >> 4;
                            const int ks = (kb == p.kb1 - 1) ? tail1 : ((kb == kbt - 1) ? tail2 : 4);
                            for (int k = 0; k < ks; ++k)
                                umma_bf16(d_tmem, hi | (uint64_t)(a16 + 2 * k), hi | (uint64_t)(b16 + 2 * k), idesc1, (kb | k) ? 1u : 0u);
                        }
                    }
                    umma_commit(g1_full(gs));
                    if (++gs == p.g1_stages) { gs = 0; gph ^= 1u; }
                }
                umma_commit(x_empty(xb));
                if (++xb == p.x_bufs) { xb = 0; xph ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ================= depthwise warps: the items of all chunks, in order, dealt round-robin =================
        const uint32_t d_mask = ~p.p_mask & ((1u << kCwCount) - 1u);          // compute warps without pw1-epilogue work
        const int n_dw = __popc(d_mask);
        const int d_bufs = p.d_bufs;
        const int my_units = (int)blockIdx.x < p.units ? (p.units - 1 - (int)blockIdx.x) / grid + 1 : 0;
        const int n_total = my_units * p.cpu;
        const int e_stages = p.e_stages;
        const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
        const uint32_t w_lane = (uint32_t)p.off_dww + (uint32_t)(lane * 8), b_lane = (uint32_t)p.off_dwb + (uint32_t)(lane * 8);
        int n = 0, i = __popc(d_mask & ((1u << (warp - kFirstCw)) - 1u));
        int s = 0; uint32_t ph = 0;
        int ds = 0; uint32_t dph = 0;
        for (;;) {
            while (i >= kItems) {
                i -= kItems; ++n;
                if (++s == e_stages) { s = 0; ph ^= 1u; }
                if (++ds == d_bufs) { ds = 0; dph ^= 1u; }
            }
            if (n >= n_total) break;
            const uint32_t st = s_e0 + (uint32_t)s * stage_bytes;
            const uint32_t d_base = s_d0 + (uint32_t)ds * kDTile;
            mbar_wait_sleep(e_full(s), ph);                                   // halo tile + depthwise weights of chunk n
            mbar_wait_sleep(d_empty(ds), dph ^ 1u);                           // the readers of the buffer's previous tile are done
            if constexpr (kBlock) {
                const int bx4 = (i & 3) * 4, by4 = (i >> 2) * 4;
                dw_block_k3<PITCH, TW>(st + (uint32_t)((by4 * PITCH + bx4) * 128), st + w_lane, st + b_lane,
                                       d_base + (uint32_t)((by4 * TW + bx4) * 128), lane, bx4 & 7);
            } else {
                const int rb = i / kStripsX, x0 = (i - rb * kStripsX) * 8, oy0 = rb * NR;
                const uint32_t e_item = st + (uint32_t)((oy0 * PITCH + x0) * 128), d_item = d_base + (uint32_t)((oy0 * TW + x0) * 128);
                if (TH % NR != 0 && oy0 + NR > TH) dw_rows<K, PITCH, TW, 1>(e_item, st + w_lane, st + b_lane, d_item, lane);
                else dw_rows<K, PITCH, TW, NR>(e_item, st + w_lane, st + b_lane, d_item, lane);
            }
            fence_proxy_async_smem();                                         // generic-proxy writes of d -> async proxy (UMMA / TMA store)
            __syncwarp();
            if (lane == 0) { mbar_arrive(d_full(ds)); mbar_arrive(e_empty(s)); }
            i += n_dw;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1 && p.mode >= 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace
}  // namespace yms

using namespace yms;

struct yms_ms_plan {
    CUtensorMap tm_e, tm_dww, tm_dwb, tm_w2, tm_y, tm_x, tm_x2, tm_w1;
    MsParams kp;
    int grid, threads;
    size_t smem;
    double flops, bytes;
};

namespace yms {
namespace {

int encode_plain(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* addr, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, const char* what) {
    auto fn = get_encode();
    if (!fn) return fail(YMS_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
    uint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(addr), dims, strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(YMS_E_DRIVER, "cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r);
    return 0;
}

template <int K> cudaError_t set_attr(int geom) {
    return geom ? cudaFuncSetAttribute(ms_layer_kernel<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit)
                : cudaFuncSetAttribute(ms_layer_kernel<K, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
}

}  // namespace
}  // namespace yms

extern "C" int yms_ms_plan_create(const yms_ms_params* q, yms_ms_plan** out) {
    if (!q || !out) return fail(YMS_E_ARG, "ms: null argument");
    *out = nullptr;
    const int k = q->ksize;
    if (k != 3 && k != 5 && k != 7 && k != 9) return fail(YMS_E_UNSUPPORTED, "ms: ksize must be 3, 5, 7 or 9");
    if (q->mode < 0 || q->mode > 2) return fail(YMS_E_ARG, "ms: mode must be 0, 1 or 2");
    if (q->batch <= 0 || q->h <= 0 || q->w <= 0 || q->e_ch <= 0 || (q->e_ch % 8)) return fail(YMS_E_ARG, "ms: bad sizes (channels % 8 == 0)");
    auto al16 = [](const void* p) { return p && ((uintptr_t)p & 15) == 0; };
    if (!al16(q->dw_weight) || !al16(q->dw_bias) || !al16(q->y) || (q->y_pixel_stride % 8)) return fail(YMS_E_ARG, "ms: null or misaligned pointer");
    if (q->mode != 2 && (!al16(q->e) || (q->e_pixel_stride % 8) || q->e_pixel_stride < q->e_ch)) return fail(YMS_E_ARG, "ms: bad expanded input");
    if (q->mode >= 1) {
        if (q->c_out <= 0 || (q->c_out % 16) || q->c_out > 256) return fail(YMS_E_UNSUPPORTED, "ms: c_out must be a multiple of 16, <= 256");
        if (!al16(q->w2) || !q->bias2 || q->y_pixel_stride < q->c_out) return fail(YMS_E_ARG, "ms: bad pw2 operands");
    } else if (q->y_pixel_stride < q->e_ch) return fail(YMS_E_ARG, "ms: bad output pixel stride");
    if (q->mode == 2) {
        if (q->c_in <= 0 || (q->c_in % 8) || q->c_in2 < 0 || (q->c_in2 % 8)) return fail(YMS_E_ARG, "ms: bad pw1 channels");
        if (!al16(q->x) || (q->x_pixel_stride % 8) || q->x_pixel_stride < q->c_in || !al16(q->w1) || !q->bias1) return fail(YMS_E_ARG, "ms: bad pw1 operands");
        if (q->c_in2 && (!al16(q->x2) || (q->x2_pixel_stride % 8) || q->x2_pixel_stride < q->c_in2)) return fail(YMS_E_ARG, "ms: bad second pw1 source");
        if (q->e_ch > 1024) return fail(YMS_E_UNSUPPORTED, "ms: fused pw1 needs e_ch <= 1024");
    }
    yms_ms_plan* pl = new (std::nothrow) yms_ms_plan();
    if (!pl) return fail(YMS_E_ARG, "ms: out of host memory");
    MsParams& kp = pl->kp;
    // tile geometry: the one that wastes fewer pixels (ties: 16 x 8); the other one when the preferred one does not fit
    const double u0 = (double)q->w * q->h / ((double)ceil_div(q->w, 16) * ceil_div(q->h, 8) * 128.0);
    const double u1 = (double)q->w * q->h / ((double)ceil_div(q->w, 24) * ceil_div(q->h, 5) * 120.0);
    const int preferred = u1 > u0 * 1.1 ? 1 : 0;
    int kTW = 16, kTH = 8, hwx = 0, hwy = 0;
    auto layout = [&](int geom) -> int {
    memset(&kp, 0, sizeof(kp));
    kp.mode = q->mode; kp.ksize = k; kp.geom = geom;
    kTW = geom ? 24 : 16; kTH = geom ? 5 : 8;
    const int kPitchB = halo_pitch(k, geom) * 128;
    kp.tiles_x = ceil_div(q->w, kTW); kp.tiles_y = ceil_div(q->h, kTH); kp.batch = q->batch;
    const long long total = (long long)kp.tiles_x * kp.tiles_y * q->batch;
    if (total * ceil_div(q->e_ch, 64) > 0x3fffffffLL) return fail(YMS_E_UNSUPPORTED, "ms: too many tiles");
    kp.total_tiles = (int)total;
    kp.mg_tiles_x = fast_div_magic(kp.tiles_x); kp.mg_tiles_y = fast_div_magic(kp.tiles_y);
    kp.e_ch = q->e_ch; kp.n_chunks = ceil_div(q->e_ch, 64);
    kp.tail_ksteps = ((q->e_ch - (kp.n_chunks - 1) * 64) + 15) >> 4;
    kp.c_out = q->mode >= 1 ? q->c_out : 0; kp.block_n = kp.c_out; kp.act2 = q->act2 ? 1 : 0;
    kp.bias2 = q->bias2; kp.bias1 = q->bias1;
    kp.img_w = q->w; kp.img_h = q->h;
    hwx = kTW + k - 1; hwy = kTH + k - 1;
    kp.hp = hwx * hwy; kp.mt = ceil_div(kp.hp, 128);
    kp.c_in1 = q->c_in; kp.c_in2 = q->c_in2;
    kp.kb1 = q->mode == 2 ? ceil_div(q->c_in, 64) : 0; kp.kb2 = q->mode == 2 ? ceil_div(q->c_in2, 64) : 0;
    kp.w1_tile_bytes = 64 * 128;
    kp.w2_tile_bytes = kp.block_n * 128;
    // one e stage: halo rows | depthwise weights [k*k][64] f32 | bias [64] f32 | (streamed pw2 weight chunk)
    const int e_rows = hwy * kPitchB;
    kp.off_dww = e_rows; kp.off_dwb = e_rows + k * k * 256;
    int stage = kp.off_dwb + 256;
    stage = (stage + 1023) & ~1023;
    kp.off_w2 = stage;
    // TMEM: pw2 accumulators (2 stages) then, in mode 2, the pw1 accumulators (2 stages x mt x 64 columns)
    int cols = 0;
    if (q->mode >= 1) {
        kp.acc_stride = kp.block_n <= 32 ? 32 : (kp.block_n <= 64 ? 64 : (kp.block_n <= 128 ? 128 : 256));
        cols = 2 * kp.acc_stride;
        if (q->mode == 2) {
            kp.g1_base = cols; kp.g1_stride = kp.mt * 64;
            kp.g1_stages = cols + 3 * kp.g1_stride <= 512 ? 3 : 2;
            cols += kp.g1_stages * kp.g1_stride;
        }
        if (cols > 512) return fail(YMS_E_UNSUPPORTED, "ms: accumulators exceed the 512 TMEM columns");
        int pow2 = 32; while (pow2 < cols) pow2 <<= 1;
        kp.tmem_cols = pow2;
    }
    kp.bias_pad = q->mode == 2 ? 256 + kp.n_chunks * 64 : 256;
    const int kbt = kp.kb1 + kp.kb2;
    kp.x_kb_stride = (kp.hp * 128 + 1023) & ~1023;
    kp.x_buf_bytes = kbt * kp.x_kb_stride;
    kp.x_bufs = q->mode == 2 ? 2 : 0;
    int x_bytes = kp.x_bufs * kp.x_buf_bytes;
    const int w1_bytes = q->mode == 2 ? kp.n_chunks * kbt * kp.w1_tile_bytes : 0;
    const int out_bytes = q->mode >= 1 ? kDTile : 0;
    const int tail_bytes = kp.bias_pad * 4 + 32 * 8 + 16;
    int fixed = x_bytes + 2 * kDTile + out_bytes + w1_bytes + tail_bytes + 1024 /* alignment slack */;
    const int w2_all = kp.n_chunks * kp.w2_tile_bytes;
    // mode 2: a second input-halo buffer lets the next tile's input land while this tile is processed; dropped when the
    // layer does not leave two e stages (+ resident pw2 weights) beside it
    if (q->mode == 2 && kSmemLimit - fixed - w2_all < 2 * stage) { kp.x_bufs = 1; x_bytes = kp.x_buf_bytes; fixed -= kp.x_buf_bytes; }
    int budget = kSmemLimit - fixed;
    kp.w2_resident = 0;
    if (q->mode >= 1) {
        // resident pw2 weights when at least 2 (mode 2) / 3 stages remain, else streamed with the halo chunk
        const int need = (q->mode == 2 ? 2 : 3) * stage;
        if (budget - w2_all >= need) { kp.w2_resident = 1; budget -= w2_all; }
        else stage += kp.w2_tile_bytes;
    }
    kp.stage_bytes = stage;
    // two e stages and two d buffers are the minimum; what is left goes, in this order, to a third d buffer (a depthwise warp that
    // finishes its item of chunk n early moves on to chunk n+2 without waiting for the consumers of chunk n), a third e stage, ...
    int stages = 2, dbufs = 2;
    int rem = budget - 2 * stage;
    if (rem < 0) return fail(YMS_E_UNSUPPORTED, "ms: the layer does not fit in shared memory (use a lower fusion mode)");
    auto take = [&](int bytes) { if (rem >= bytes) { rem -= bytes; return true; } return false; };
    if (take(kDTile)) dbufs = 3;
    if (take(stage)) stages = 3;
    if (dbufs == 3 && take(kDTile)) dbufs = 4;
    if (stages == 3 && take(stage)) stages = 4;
    kp.e_stages = stages;
    kp.d_bufs = dbufs;
    kp.n_items = items_per_chunk(k, geom);
    // mode 2: compute warp w (TMEM lane quadrant (kFirstCw + w) & 3, M-tile parity w >> 2) is a pw1-epilogue warp iff one of its
    // M tiles holds halo pixels in its quadrant (18 x 10 = 180 halo pixels: 6 of the 8 candidates); the rest join the depthwise warps
    kp.p_mask = 0; kp.n_p = 0;
    if (q->mode == 2)
        for (int w = 0; w < kPWarps; ++w) {
            const int wq = (kFirstCw + w) & 3;
            bool any = false;
            for (int m = w >> 2; m < kp.mt; m += 2) any |= m * 128 + wq * 32 < kp.hp;
            if (any) { kp.p_mask |= 1u << w; ++kp.n_p; }
        }
    kp.units = q->mode == 0 ? kp.total_tiles * kp.n_chunks : kp.total_tiles;
    kp.cpu = q->mode == 0 ? 1 : kp.n_chunks;
    kp.stage_tx = (uint32_t)((q->mode != 2 ? hwx * hwy * 128 : 0) + k * k * 256 + 256 + ((q->mode >= 1 && !kp.w2_resident) ? kp.w2_tile_bytes : 0));
    kp.so_x = 0;
    kp.so_e = x_bytes;
    kp.so_d = kp.so_e + stages * stage;
    kp.so_w2 = kp.so_d + dbufs * kDTile;
    kp.so_out = kp.so_w2 + (kp.w2_resident ? w2_all : 0);
    kp.so_w1 = kp.so_out + out_bytes;
    kp.so_bias = kp.so_w1 + w1_bytes;
    pl->smem = (size_t)kp.so_bias + tail_bytes + 1024;
    if (pl->smem > (size_t)kSmemLimit) return fail(YMS_E_UNSUPPORTED, "ms: shared-memory budget exceeded");
    return 0;
    };
    int lrc = layout(preferred);
    if (lrc == YMS_E_UNSUPPORTED) lrc = layout(1 - preferred);
    if (lrc) { delete pl; return lrc; }
    pl->grid = kp.units < kNumSMs ? kp.units : kNumSMs;
    pl->threads = kThreads;

    int rc = 0;
    const uint64_t W = (uint64_t)q->w, H = (uint64_t)q->h, N = (uint64_t)q->batch;
    auto act_map = [&](CUtensorMap* m, const void* ptr, int c, int64_t ps, int bx, int by, const char* what) {
        uint64_t dims[4] = {(uint64_t)c, W, H, N};
        uint64_t strides[3] = {(uint64_t)ps * 2, (uint64_t)ps * 2 * W, (uint64_t)ps * 2 * W * H};
        uint32_t box[4] = {64, (uint32_t)bx, (uint32_t)by, 1};
        uint32_t es[4] = {1, 1, 1, 1};
        return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, what);
    };
    if (q->mode != 2) rc = act_map(&pl->tm_e, q->e, q->e_ch, q->e_pixel_stride, hwx, halo_pitch(k, kp.geom) == hwx ? hwy : 1, "ms e");
    else {
        rc = act_map(&pl->tm_x, q->x, q->c_in, q->x_pixel_stride, hwx, hwy, "ms x");
        if (!rc && q->c_in2) rc = act_map(&pl->tm_x2, q->x2, q->c_in2, q->x2_pixel_stride, hwx, hwy, "ms x2");
        if (!rc) {
            const uint64_t K1 = (uint64_t)(q->c_in + q->c_in2);
            uint64_t dims[3] = {K1, (uint64_t)q->e_ch, 1};
            uint64_t strides[2] = {K1 * 2, K1 * 2 * (uint64_t)q->e_ch};
            uint32_t box[3] = {64, 64, 1};
            uint32_t es[3] = {1, 1, 1};
            rc = encode_map(&pl->tm_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->w1, dims, strides, box, es, "ms w1");
        }
    }
    if (q->mode != 2) { pl->tm_x = pl->tm_e; pl->tm_w1 = pl->tm_e; }
    if (q->mode != 2 || !q->c_in2) pl->tm_x2 = pl->tm_x;
    if (q->mode == 2) pl->tm_e = pl->tm_x;
    if (!rc) rc = act_map(&pl->tm_y, q->y, q->mode >= 1 ? q->c_out : q->e_ch, q->y_pixel_stride, kTW, kTH, "ms y");
    if (!rc) {
        uint64_t dims[2] = {(uint64_t)q->e_ch, (uint64_t)(k * k)};
        uint64_t strides[1] = {(uint64_t)q->e_ch * 4};
        uint32_t box[2] = {64, (uint32_t)(k * k)};
        rc = encode_plain(&pl->tm_dww, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, q->dw_weight, dims, strides, box, "ms dw weight");
    }
    if (!rc) {
        uint64_t dims[2] = {(uint64_t)q->e_ch, 1};
        uint64_t strides[1] = {(uint64_t)q->e_ch * 4};
        uint32_t box[2] = {64, 1};
        rc = encode_plain(&pl->tm_dwb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, q->dw_bias, dims, strides, box, "ms dw bias");
    }
    if (!rc && q->mode >= 1) {
        uint64_t dims[3] = {(uint64_t)q->e_ch, (uint64_t)q->c_out, 1};
        uint64_t strides[2] = {(uint64_t)q->e_ch * 2, (uint64_t)q->e_ch * 2 * (uint64_t)q->c_out};
        uint32_t box[3] = {64, (uint32_t)kp.block_n, 1};
        uint32_t es[3] = {1, 1, 1};
        rc = encode_map(&pl->tm_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, q->w2, dims, strides, box, es, "ms w2");
    } else if (!rc) pl->tm_w2 = pl->tm_y;
    if (rc) { delete pl; return rc; }

    const double px = (double)q->batch * q->h * q->w;
    pl->flops = 2.0 * px * q->e_ch * k * k;
    pl->bytes = (double)(k * k + 1) * q->e_ch * 4.0;
    if (q->mode == 0) pl->bytes += 2.0 * px * q->e_ch * 2.0;
    if (q->mode == 1) { pl->flops += 2.0 * px * q->e_ch * q->c_out; pl->bytes += 2.0 * px * (q->e_ch + q->c_out) + 2.0 * q->e_ch * q->c_out; }
    if (q->mode == 2) {
        const double K1 = q->c_in + q->c_in2;
        pl->flops += 2.0 * px * q->e_ch * (q->c_out + K1);
        pl->bytes += 2.0 * px * (K1 + q->c_out) + 2.0 * q->e_ch * (q->c_out + K1);
    }

    cudaError_t e = cudaSuccess;        // per call: the attribute is per device, and plan creation is not a hot path
    switch (k) {
        case 3: e = set_attr<3>(kp.geom); break;
        case 5: e = set_attr<5>(kp.geom); break;
        case 7: e = set_attr<7>(kp.geom); break;
        default: e = set_attr<9>(kp.geom); break;
    }
    if (e != cudaSuccess) { delete pl; return fail((int)e, "ms: smem attribute: %s", cudaGetErrorString(e)); }
    *out = pl;
    return 0;
}

extern "C" int yms_ms_plan_run(const yms_ms_plan* pl, void* stream) {
    if (!pl) return fail(YMS_E_ARG, "ms: null plan");
    cudaError_t le;
    cudaStream_t st = (cudaStream_t)stream;
#define YMS_MS_LAUNCH(K) le = pl->kp.geom ? launch_pdl(ms_layer_kernel<K, 1>, pl->grid, pl->threads, pl->smem, st, pl->tm_e, pl->tm_dww, pl->tm_dwb, \
                                                      pl->tm_w2, pl->tm_y, pl->tm_x, pl->tm_x2, pl->tm_w1, pl->kp) \
                                         : launch_pdl(ms_layer_kernel<K, 0>, pl->grid, pl->threads, pl->smem, st, pl->tm_e, pl->tm_dww, pl->tm_dwb, \
                                                      pl->tm_w2, pl->tm_y, pl->tm_x, pl->tm_x2, pl->tm_w1, pl->kp)
    switch (pl->kp.ksize) {
        case 3: YMS_MS_LAUNCH(3); break;
        case 5: YMS_MS_LAUNCH(5); break;
        case 7: YMS_MS_LAUNCH(7); break;
        default: YMS_MS_LAUNCH(9); break;
    }
#undef YMS_MS_LAUNCH
    if (le != cudaSuccess) return fail((int)le, "ms_layer_kernel launch: %s", cudaGetErrorString(le));
    return check_launch("ms_layer_kernel");
}

extern "C" int yms_ms_plan_destroy(yms_ms_plan* pl) {
    delete pl;
    return 0;
}

extern "C" int yms_ms_plan_cost(const yms_ms_plan* pl, double* flops, double* bytes) {
    if (!pl) return fail(YMS_E_ARG, "ms: null plan");
    if (flops) *flops = pl->flops;
    if (bytes) *bytes = pl->bytes;
    return 0;
}

/* Conv(c, c, k, 1, k//2, groups=c) of components.py:69-77: mode 0 of the kernel above (tensor maps are encoded per call; inside a
 * captured CUDA graph that cost is paid once). */
extern "C" int yms_dwconv(const void* x, int64_t xps, int batch, int h, int w, int channels, int ksize,
                          const float* weight, const float* bias, void* y, int64_t yps, void* stream) {
    yms_ms_params q;
    memset(&q, 0, sizeof(q));
    q.mode = 0; q.batch = batch; q.h = h; q.w = w; q.ksize = ksize; q.e_ch = channels;
    q.e = x; q.e_pixel_stride = xps; q.y = y; q.y_pixel_stride = yps;
    q.dw_weight = weight; q.dw_bias = bias;
    yms_ms_plan* pl = nullptr;
    int rc = yms_ms_plan_create(&q, &pl);
    if (rc) return rc;
    rc = yms_ms_plan_run(pl, stream);
    delete pl;
    return rc;
}
