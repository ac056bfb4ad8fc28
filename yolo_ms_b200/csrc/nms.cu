// Batched class-aware NMS for the YOLO post-process: ONE kernel launch per batch.
//
// Replaces yolov8/tools/test.py:181-218 (confidence filter + `for c in unique(cls): nms(...)`
// around torchvision.ops.nms) of the reference.  Results are bit-identical to that path:
//   * candidates: score > conf (strict, fp32)
//   * order: label ascending, score descending, ties by lower index (== stable sort per class)
//   * suppression iff (double)IoU > iou_thr, IoU = inter / (area_i + area_j - inter) in fp32
//     with every operation individually rounded (no FMA contraction), NaN never suppresses.
//
// Work decomposition: classes are independent, so every image is handled by G CTAs (host picks
// G ~ 148 / batch), CTA g owning the contiguous class range [g*nc/G, (g+1)*nc/G).  Per CTA:
//   A  compact the candidates of its classes into 64-bit keys [label:11 | ~score:32 | index:20]
//   A  ... scattered straight into their class segment (the class histogram gives the offsets: counting sort on the label)
//   B  per-segment sort by (score desc, index) in shared memory: one warp per small class, the CTA for big ones
//      (YMS_NMS_SORT=bitonic / key sets beyond shared memory: one bitonic sort of all keys, size = next pow2 of the count)
//   C  class segment table
//   D  greedy suppression.  Fast path (<= 8192 candidates: keys AND boxes in shared memory):
//      segments are cut into chunks of 32 sorted boxes, chunks are dealt round-robin to the 32
//      warps and processed as a software pipeline: a chunk streams the already-kept boxes of its
//      class past its lanes (shfl broadcast) WHILE the preceding chunks are still being resolved,
//      and only then resolves itself in-register; per-class progress words in shared memory
//      publish (finished chunks, kept count); __ballot_sync/__popc compact survivors in place.
//      General path (any size, keys possibly in a global workspace): one warp per class segment.
//   E  per-CTA keep list -> staging; the LAST CTA of an image (atomic ticket) concatenates the G
//      lists in class order into the final keep list.
#include "common.cuh"

#include <type_traits>

#include <stdlib.h>

namespace yms {
extern long long* g_prof_buf;
namespace {

#ifdef YMS_PROF
#define NMS_STAMP(i) do { if (a.prof && threadIdx.x == 0) a.prof[16 * blockIdx.x + (i)] = clock64(); } while (0)
#else
#define NMS_STAMP(i) do { } while (0)
#endif

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kSortTile = 16384;           // u64 keys that fit the 128 KB key region
constexpr int kFastCap = 8192;             // fast path: 64 KB keys + 128 KB boxes
constexpr int kMaxClasses = 2047;
constexpr int kWarpSortMax = 256;          // class segments up to this many keys are sorted by one warp, larger ones by the CTA
constexpr int kMaxGroups = 16;
constexpr unsigned long long kInvalidKey = ~0ull;
constexpr unsigned long long kIdxMask = (1ull << 20) - 1;
constexpr size_t kKeyRegionBytes = (size_t)kSortTile * 8 + (size_t)kFastCap * 8;   // 192 KB: keys | boxes

__device__ __forceinline__ uint32_t desc_score_bits(float s) {
    s = s + 0.0f;                               // -0 -> +0 (torch's sort treats them as equal)
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone ascending map
    return ~u;                                  // descending
}

// max/min exactly as std::max / std::min (torchvision CPU kernel) incl. NaN behaviour
__device__ __forceinline__ float max_std(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float min_std(float a, float b) { return (b < a) ? b : a; }

__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// true iff box j must be suppressed by kept box i
__device__ __forceinline__ bool suppresses(const float4& bi, float ai, const float4& bj, float aj, float thr_f) {
    float xx1 = max_std(bi.x, bj.x);
    float yy1 = max_std(bi.y, bj.y);
    float xx2 = min_std(bi.z, bj.z);
    float yy2 = min_std(bi.w, bj.w);
    float w = max_std(0.0f, __fsub_rn(xx2, xx1));
    float h = max_std(0.0f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    if (!(inter > 0.0f)) return false;          // ovr is 0 or NaN: never > thr (thr >= 0)
    float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    // Division-free decision whenever it is provably the same as the reference's rounded quotient: with
    // q = inter / uni (real), fl(q) > thr_f is certain for inter >= thr_f*uni*(1 + 2e-6) and impossible for
    // inter <= thr_f*uni*(1 - 2e-6) (two fp32 products: <= 1.2e-7 relative error; the quotient's rounding: 6e-8;
    // the magnitude guards keep every value normal).  Only the 4e-6-wide band in between pays for __fdiv_rn.
    if (inter > 1e-15f && uni > 1e-15f && uni < 1e15f) {
        const float t = __fmul_rn(thr_f, uni);
        if (inter >= __fmul_rn(t, 1.000002f)) return true;
        if (inter <= __fmul_rn(t, 0.999998f)) return false;
    }
    float ovr = __fdiv_rn(inter, uni);
    // (double)ovr > thr  <=>  ovr > thr_f with thr_f = largest float <= thr (host computes it)
    return ovr > thr_f;
}

// bitonic compare-exchange steps j = j_first, j_first/2, ..., 1 of merge size k on a tile of
// `tile` keys in shared memory whose first element has global index gbase.
__device__ void bitonic_tile_steps(unsigned long long* s, int tile, int gbase, int k, int j_first) {
    for (int j = j_first; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (tile >> 1); t += blockDim.x) {
            int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            int l = i | j;
            bool up = (((gbase + i) & k) == 0);
            unsigned long long a = s[i], b = s[l];
            if ((a > b) == up) { s[i] = b; s[l] = a; }
        }
        __syncthreads();
    }
}

// Ascending sort of one class segment s[0, len) in shared memory by `nthreads` cooperating threads (a warp, or the whole CTA
// when kBlock).  Normalised bitonic network: every merge starts with a MIRROR step (i <-> k-1-i inside a k block) followed
// by half-cleaners, so every compare-exchange puts the minimum at the LOWER index; positions >= len then behave as +inf
// without being stored, and a segment of any length is sorted in place (no power-of-two padding region).
template <bool kBlock>
__device__ __forceinline__ void segment_sort(unsigned long long* s, int len, int t0, int nthreads) {
    if (len < 2) return;                                   // uniform over the cooperating threads
    int P = 2;
    while (P < len) P <<= 1;
    const int half_pairs = P >> 1;
    for (int k = 2, lg = 0; k <= P; k <<= 1, ++lg) {        // lg = log2(k / 2)
        for (int t = t0; t < half_pairs; t += nthreads) {
            const int blk = t >> lg, off = t & ((k >> 1) - 1);
            const int i = blk * k + off, l = blk * k + (k - 1 - off);
            if (l < len) {
                const unsigned long long x = s[i], y = s[l];
                if (x > y) { s[i] = y; s[l] = x; }
            }
        }
        if (kBlock) __syncthreads(); else __syncwarp();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = t0; t < half_pairs; t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                if (l < len) {
                    const unsigned long long x = s[i], y = s[l];
                    if (x > y) { s[i] = y; s[l] = x; }
                }
            }
            if (kBlock) __syncthreads(); else __syncwarp();
        }
    }
}

// Apply kept boxes kb[from, to) (compact, score order) to the lane's box; returns updated flag.
template <typename BoxAt>
__device__ __forceinline__ bool apply_kept(BoxAt box_at, int from, int to, const float4& bj, float aj, bool removed,
                                           float thr_f, int lane) {
    for (int g = from; g < to; g += 32) {
        if (__all_sync(0xffffffffu, removed)) break;
        const int cnt = min(32, to - g);
        float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < cnt) bp = box_at(g + lane);
        const float ap = box_area(bp);
        for (int p = 0; p < cnt; ++p) {
            float4 bi;
            bi.x = __shfl_sync(0xffffffffu, bp.x, p);
            bi.y = __shfl_sync(0xffffffffu, bp.y, p);
            bi.z = __shfl_sync(0xffffffffu, bp.z, p);
            bi.w = __shfl_sync(0xffffffffu, bp.w, p);
            float ai = __shfl_sync(0xffffffffu, ap, p);
            if (!removed && suppresses(bi, ai, bj, aj, thr_f)) removed = true;
        }
    }
    return removed;
}

// Register + shuffle flavour of the same normalised bitonic network for ONE warp and segments of at most 32 * E keys
// (scripts/ubench/seg_sort.cu, measured on the B200: 10.2 kcycles against 28.8 kcycles for the shared-memory network on the
// bench workload's 24 segments per CTA -- the shared-memory version is bound by its 2 x 8-byte accesses per compare-exchange).
// Element index = lane * E + r.  Partners below E stay in the lane; a half-cleaner of stride j >= E pairs lane with
// lane ^ (j / E); the mirror step of merge size k > E pairs (lane, r) with (lane ^ (k / E - 1), E - 1 - r).  Positions >= len
// hold +inf and never move.
__device__ __forceinline__ void cmpx64(unsigned long long& x, unsigned long long& y) { if (x > y) { const unsigned long long t = x; x = y; y = t; } }

template <int E>
__device__ __forceinline__ void warp_sort_regs(unsigned long long* s, int len, int lane) {
    unsigned long long v[E];
    #pragma unroll
    for (int r = 0; r < E; ++r) { const int i = lane * E + r; v[r] = (i < len) ? s[i] : ~0ull; }
    constexpr int P = 32 * E;
    #pragma unroll
    for (int k = 2; k <= P; k <<= 1) {
        if (k <= E) {
            #pragma unroll
            for (int r = 0; r < E; ++r) { const int q = r ^ (k - 1); if (q > r) cmpx64(v[r], v[q]); }
        } else {
            const int mm = k / E - 1;
            const bool lower = (lane & ((mm + 1) >> 1)) == 0;
            unsigned long long o[E];
            #pragma unroll
            for (int r = 0; r < E; ++r) o[r] = __shfl_xor_sync(0xffffffffu, v[E - 1 - r], mm);
            #pragma unroll
            for (int r = 0; r < E; ++r) v[r] = lower ? (v[r] < o[r] ? v[r] : o[r]) : (v[r] > o[r] ? v[r] : o[r]);
        }
        #pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j < E) {
                #pragma unroll
                for (int r = 0; r < E; ++r) if ((r & j) == 0) cmpx64(v[r], v[r | j]);
            } else {
                const int m = j / E;
                const bool lower = (lane & m) == 0;
                #pragma unroll
                for (int r = 0; r < E; ++r) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], m);
                    v[r] = lower ? (v[r] < o ? v[r] : o) : (v[r] > o ? v[r] : o);
                }
            }
        }
    }
    __syncwarp();
    #pragma unroll
    for (int r = 0; r < E; ++r) { const int i = lane * E + r; if (i < len) s[i] = v[r]; }
    __syncwarp();
}

__device__ __forceinline__ void warp_sort_segment(unsigned long long* s, int len, int lane) {     // len <= kWarpSortMax
    if (len < 2) return;
    if (len <= 32) warp_sort_regs<1>(s, len, lane);
    else if (len <= 64) warp_sort_regs<2>(s, len, lane);
    else if (len <= 128) warp_sort_regs<4>(s, len, lane);
    else warp_sort_regs<8>(s, len, lane);
}

// Resolve one chunk (32 boxes in score order held by the lanes); returns the survivor mask.
__device__ __forceinline__ unsigned resolve_chunk(const float4& bj, float aj, bool& removed, float thr_f, int lane) {
    unsigned alive = ~__ballot_sync(0xffffffffu, removed);
    for (int i = 0; i < 32; ++i) {
        if (!((alive >> i) & 1u)) continue;            // warp-uniform
        float4 bi;
        bi.x = __shfl_sync(0xffffffffu, bj.x, i);
        bi.y = __shfl_sync(0xffffffffu, bj.y, i);
        bi.z = __shfl_sync(0xffffffffu, bj.z, i);
        bi.w = __shfl_sync(0xffffffffu, bj.w, i);
        float ai = __shfl_sync(0xffffffffu, aj, i);
        bool hit = (lane > i) && !removed && suppresses(bi, ai, bj, aj, thr_f);
        if (hit) removed = true;
        alive &= ~__ballot_sync(0xffffffffu, hit);
    }
    return alive;
}

// ---- IoU bitmask path (fast path when the masks fit in shared memory) ------------------------------------------
// A class segment of n sorted boxes is cut into W = ceil(n/32) blocks; tile (k, w >= k) of its strictly-upper-
// triangular suppression matrix is 32 words: word l = the 32 bits "row 32k+l suppresses column 32w+j".
//   phase 1 (all 32 warps, tiles of all classes dealt round-robin): every pair test is independent -- no serial
//            chain through the greedy order;
//   phase 2 (one warp per class): the greedy order is replayed on the words alone -- per block a 32-step shuffle
//            sweep over the diagonal tile decides which rows survive, the words of the survivors against later
//            columns are OR-reduced (__reduce_or_sync) into `removed` (lane w holds word w, so W <= 32).
// Same pair predicate and same order as the reference => bit-identical keep lists.

// Pair predicate for FINITE boxes (the bitmask path is only taken when every box of the CTA is finite): fmaxf/fminf
// equal std::max/std::min on finite values up to the sign of zero (which cannot make `inter > 0` true), every
// operation is individually rounded like the reference's, and the decision is branch-free except for the 4e-6-wide
// band around the threshold that needs the exactly rounded quotient.
__device__ __forceinline__ bool suppresses_finite(const float4& bi, float ai, const float4& bj, float aj, float thr_f) {
    const float w = fmaxf(__fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    const float t = __fmul_rn(thr_f, uni);
    const bool guard = inter > 1e-15f && uni > 1e-15f && uni < 1e15f;
    const bool sure_t = guard && inter >= __fmul_rn(t, 1.000002f);
    const bool sure_f = (guard && inter <= __fmul_rn(t, 0.999998f)) || !(inter > 0.0f);
    bool hit = sure_t;
    if (!(sure_t || sure_f)) hit = __fdiv_rn(inter, uni) > thr_f;
    return hit;
}

// Tile of the suppression matrix with a two-stage conservative pre-filter.  For proper boxes (w, h >= 0; checked per
// CTA) IoU > t implies  x-overlap > t * max(w_i, w_j)  AND  y-overlap > t * max(h_i, h_j)  (inter = ix*iy <= ix*h_i and
// union >= w_i*h_i, symmetrically for j and for y).  Both tests use t * 0.999, far more slack than fp32 rounding needs,
// so they never reject a pair the exact predicate accepts; they are warp-voted, so that the ~40-instruction exact
// predicate only runs for the few column boxes that at least one of the 32 row boxes can possibly suppress.
template <bool kDiag>
__device__ __forceinline__ void nms_mask_tile(const float4* sb, int n, int k, int w, unsigned* out, float thr_f, int lane) {
    const int r = 32 * k + lane;
    const bool rv = r < n;
    const float4 bi = rv ? sb[r] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ai = box_area(bi);
    const float wi = __fsub_rn(bi.z, bi.x), hi = __fsub_rn(bi.w, bi.y);
    const float tq = thr_f * 0.999f;
    const int jn = min(32, n - 32 * w);
    unsigned word = 0u;
    #pragma unroll 4
    for (int jj = 0; jj < jn; ++jj) {                                            // warp-uniform trip count, broadcast reads
        const float4 bj = sb[32 * w + jj];
        const float ix = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
        bool cand = rv && (!kDiag || lane < jj) && ix > tq * fmaxf(wi, __fsub_rn(bj.z, bj.x));
        if (!__any_sync(0xffffffffu, cand)) continue;
        const float iy = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
        cand = cand && iy > tq * fmaxf(hi, __fsub_rn(bj.w, bj.y));
        if (!__any_sync(0xffffffffu, cand)) continue;
        const bool hit = cand && suppresses_finite(bi, ai, bj, box_area(bj), thr_f);
        word |= hit ? (1u << jj) : 0u;
    }
    out[lane] = word;
}

// Greedy replay of row blocks [kb, ke) of one class on the mask words alone (one warp).  `tiles` holds, row block after row
// block, the W-k tiles (k, k..W-1) of the round; remw[w] (shared memory, owned by this warp) accumulates the removed bits.
__device__ __forceinline__ void nms_mask_sweep_rows(const unsigned* tiles, unsigned* remw, int n, int kb, int ke, int lane) {
    const int W = (n + 31) >> 5;
    for (int k = kb; k < ke; ++k) {
        const unsigned diag = tiles[lane];
        unsigned rem = remw[k];
        #pragma unroll
        for (int i = 0; i < 32; ++i) {
            const unsigned mi = __shfl_sync(0xffffffffu, diag, i);
            rem |= ((rem >> i) & 1u) ? 0u : mi;                                  // row i survives -> it removes its later neighbours
        }
        const bool alive = (32 * k + lane < n) && !((rem >> lane) & 1u);
        if (lane == 0) remw[k] = rem;
        for (int w = k + 1; w < W; ++w) {
            const unsigned red = __reduce_or_sync(0xffffffffu, alive ? tiles[(w - k) * 32 + lane] : 0u);
            if (lane == 0 && red) remw[w] |= red;
        }
        __syncwarp();
        tiles += (size_t)(W - k) * 32;
    }
}

// survivor keys compacted in place (score order preserved); returns the kept count
__device__ __forceinline__ int nms_mask_compact(unsigned long long* sk, const unsigned* remw, int n, int lane) {
    const int W = (n + 31) >> 5;
    int kept = 0;
    for (int k = 0; k < W; ++k) {
        const int r = 32 * k + lane;
        const unsigned long long key = (r < n) ? sk[r] : kInvalidKey;
        const bool alive = (r < n) && !((remw[k] >> lane) & 1u);
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        __syncwarp();                                                            // the block's keys are in registers before its (lower) slots are rewritten
        if (alive) sk[kept + __popc(bal & ((1u << lane) - 1u))] = key;
        kept += __popc(bal);
    }
    return kept;
}

// ---- pipelined greedy path, shared-memory flavour ------------------------------------------------------------------
template <bool kFinite>
__device__ __forceinline__ bool pair_hit(const float4& bi, float ai, const float4& bj, float aj, float thr_f) {
    return kFinite ? suppresses_finite(bi, ai, bj, aj, thr_f) : suppresses(bi, ai, bj, aj, thr_f);
}

// Apply the kept boxes kb[from, to) of the class (compact, score order, in shared memory) to the lane's box.  The kept
// boxes are broadcast reads (no shuffles) and four of them are tested per iteration: the tests are independent, which
// gives the single warp that owns a chunk the instruction-level parallelism its dependent predicate chain lacks.
template <bool kFinite>
__device__ __forceinline__ bool apply_kept_smem(const float4* kb, int from, int to, const float4& bj, float aj, bool removed, float thr_f) {
    int g = from;
    for (; g + 4 <= to; g += 4) {
        if (__all_sync(0xffffffffu, removed)) return true;
        const float4 b0 = kb[g], b1 = kb[g + 1], b2 = kb[g + 2], b3 = kb[g + 3];
        const bool h0 = pair_hit<kFinite>(b0, box_area(b0), bj, aj, thr_f);
        const bool h1 = pair_hit<kFinite>(b1, box_area(b1), bj, aj, thr_f);
        const bool h2 = pair_hit<kFinite>(b2, box_area(b2), bj, aj, thr_f);
        const bool h3 = pair_hit<kFinite>(b3, box_area(b3), bj, aj, thr_f);
        removed = removed || h0 || h1 || h2 || h3;
    }
    for (; g < to; ++g) {
        const float4 b0 = kb[g];
        removed = removed || pair_hit<kFinite>(b0, box_area(b0), bj, aj, thr_f);
    }
    return removed;
}

// Division-free part of the finite predicate: straight-line code (no branch: four of these interleave), the decision in two
// flags.  `unc` marks the pairs that need the exactly rounded quotient (the 4e-6-wide band around the threshold, or values
// outside the magnitude guards).
struct PairFlags { bool hit, unc; float inter, uni; };
__device__ __forceinline__ PairFlags pair_flags(const float4& bi, const float4& bj, float aj, float thr_f) {
    PairFlags r;
    const float ai = box_area(bi);
    const float w = fmaxf(__fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 0.0f);
    r.inter = __fmul_rn(w, h);
    r.uni = __fsub_rn(__fadd_rn(ai, aj), r.inter);
    const float t = __fmul_rn(thr_f, r.uni);
    const bool guard = r.inter > 1e-15f && r.uni > 1e-15f && r.uni < 1e15f;
    r.hit = guard && r.inter >= __fmul_rn(t, 1.000002f);
    const bool sure_f = (guard && r.inter <= __fmul_rn(t, 0.999998f)) || !(r.inter > 0.0f);
    r.unc = !(r.hit || sure_f);
    return r;
}

// Same, with the kept boxes given as positions into the CTA's (immutable) sorted box array: kp[from, to) -> sb[kp[g]].
template <bool kFinite, typename BoxAt>
__device__ __forceinline__ bool apply_kept_pos(BoxAt sb, const unsigned short* kp, int from, int to, const float4& bj, float aj,
                                               bool removed, float thr_f) {
    int g = from;
    for (; g + 4 <= to; g += 4) {
        if (__all_sync(0xffffffffu, removed)) return true;
        const int p0 = kp[g], p1 = kp[g + 1], p2 = kp[g + 2], p3 = kp[g + 3];
        const float4 b0 = sb(p0), b1 = sb(p1), b2 = sb(p2), b3 = sb(p3);
        if (kFinite) {
            const PairFlags f0 = pair_flags(b0, bj, aj, thr_f), f1 = pair_flags(b1, bj, aj, thr_f);
            const PairFlags f2 = pair_flags(b2, bj, aj, thr_f), f3 = pair_flags(b3, bj, aj, thr_f);
            removed = removed || f0.hit || f1.hit || f2.hit || f3.hit;
            if (__any_sync(0xffffffffu, f0.unc || f1.unc || f2.unc || f3.unc)) {       // rare: some pair sits in the threshold band
                if (f0.unc && __fdiv_rn(f0.inter, f0.uni) > thr_f) removed = true;
                if (f1.unc && __fdiv_rn(f1.inter, f1.uni) > thr_f) removed = true;
                if (f2.unc && __fdiv_rn(f2.inter, f2.uni) > thr_f) removed = true;
                if (f3.unc && __fdiv_rn(f3.inter, f3.uni) > thr_f) removed = true;
            }
        } else {
            const bool h0 = pair_hit<false>(b0, box_area(b0), bj, aj, thr_f);
            const bool h1 = pair_hit<false>(b1, box_area(b1), bj, aj, thr_f);
            const bool h2 = pair_hit<false>(b2, box_area(b2), bj, aj, thr_f);
            const bool h3 = pair_hit<false>(b3, box_area(b3), bj, aj, thr_f);
            removed = removed || h0 || h1 || h2 || h3;
        }
    }
    for (; g < to; ++g) {
        const float4 b0 = sb(kp[g]);
        removed = removed || pair_hit<kFinite>(b0, box_area(b0), bj, aj, thr_f);
    }
    return removed;
}

struct NmsArgs {
    const float4* boxes; const float* scores; const int32_t* labels; const int32_t* n_valid;
    int n, num_classes, groups; float conf; float thr_f;
    int32_t* keep; int32_t* keep_count;
    unsigned long long* ws_keys;   // [B*G][pow2(n)] (only when pow2(n) > kSortTile)
    int32_t* ws_stage;             // [B*G][n]  per-CTA keep lists
    int32_t* ws_count;             // [B*G]
    unsigned int* ws_ticket;       // [B] zeroed by the host before the launch
    int n_pad_full;                // pow2(n)
    long long* prof;               // -DYMS_PROF builds: [grid][16] phase time stamps
    int mask_tile_limit;           // bitmask path only when the largest class of the CTA has at most this many 32-box blocks
    int poll_ns;                   // back-off unit of the pipelined greedy path's progress polling (library option nms_poll_ns, default 256 ns per chunk ahead)
    int seg_sort;                  // 1 (default): counting scatter by class + per-segment sorts; 0 (library option nms_sort_bitonic): one bitonic sort of all keys
    int region_bytes;              // shared-memory bytes of the key | box | kept-position (or mask tile) region in front of the class tables
};

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(NmsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);
    const int G = a.groups;
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int nc = a.num_classes;
    int* cls_start = reinterpret_cast<int*>(smem_raw + a.region_bytes);             // [nc + 1] (a CTA may own up to nc classes)
    int* cls_count = cls_start + (nc + 1);                                          // [nc + 1] kept counts, then offsets
    int* chunk_base = cls_count + (nc + 1);                                         // [nc + 1]; first: the image's class histogram
    volatile unsigned* state = reinterpret_cast<volatile unsigned*>(chunk_base + (nc + 1));   // [nc + 1] (done chunks<<16 | kept) / next row block
    unsigned* remw_all = const_cast<unsigned*>(reinterpret_cast<volatile unsigned*>(state + (nc + 1)));   // [kSortTile/32 + nc + 2] removed words
    __shared__ int s_count, s_next, s_last, s_clo, s_chi;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const int nb = a.n_valid ? min(max(a.n_valid[b], 0), n) : n;
    const float4* boxes = a.boxes + (size_t)b * n;
    const float* scores = a.scores + (size_t)b * n;
    const int32_t* labels = a.labels + (size_t)b * n;

    NMS_STAMP(0);
    if (tid == 0) { s_count = 0; s_next = 0; s_last = 0; }
    int* hist = chunk_base;
    for (int c = tid; c <= nc; c += kNmsThreads) hist[c] = 0;
    __syncthreads();

    // ---- A0: class histogram of the image's candidates; the G CTAs of an image all derive the SAME contiguous class
    // ranges from it, balanced by the pair-test cost W(W+1)/2 (W = ceil(n_c / 32)) instead of by class count ---------
    for (int i = tid; i < nb; i += kNmsThreads) {
        const float sc = scores[i]; const int lab = labels[i];
        if (sc > a.conf && lab >= 0 && lab < nc) atomicAdd(&hist[lab], 1);
    }
    __syncthreads();
    if (warp == 0) {
        // cost of a class = its share of the image's pair-test tiles + its share of the image's candidates (both matter:
        // tiles bound the suppression work, the count must stay inside the shared-memory fast path of every CTA)
        long long t_tot = 0, n_tot = 0;
        for (int c = lane; c < nc; c += 32) { const int w = (hist[c] + 31) >> 5; t_tot += (w * (w + 1)) >> 1; n_tot += hist[c]; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { t_tot += __shfl_xor_sync(0xffffffffu, t_tot, o); n_tot += __shfl_xor_sync(0xffffffffu, n_tot, o); }
        t_tot = max(t_tot, 1LL); n_tot = max(n_tot, 1LL);
        const long long total = 2 * t_tot * n_tot + nc;
        // class c belongs to group floor(cost_before(c) * G / total): monotone => contiguous ranges
        long long running = 0; int lo = nc, hi = 0, cnt = 0;
        for (int base = 0; base < nc; base += 32) {
            const int c = base + lane;
            const int w = (c < nc) ? ((hist[c] + 31) >> 5) : 0;
            const long long v = (c < nc) ? (long long)((w * (w + 1)) >> 1) * n_tot + (long long)hist[c] * t_tot + 1 : 0;
            long long incl = v;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (c < nc) {
                int grp = (int)(((running + incl - v) * G) / total);
                if (grp > G - 1) grp = G - 1;
                if (grp == g) { lo = min(lo, c); hi = max(hi, c + 1); cnt += hist[c]; }
            }
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (lane == 0) { s_clo = (hi > lo) ? lo : 0; s_chi = (hi > lo) ? hi : 0; s_count = cnt; }
    }
    __syncthreads();
    const int c_lo = s_clo, c_hi = s_chi;
    const int ncl = c_hi - c_lo;                                                    // classes owned by this CTA
    const int m = s_count;
    int n_pad = 32;
    while (n_pad < m) n_pad <<= 1;
    // fast path: keys, boxes and the kept-position list (2 bytes per candidate) of the CTA fit the region
    const bool fast = (m <= kFastCap) && ((size_t)n_pad * 8 + (size_t)m * 18 + 64 <= (size_t)a.region_bytes);
    const bool in_smem = (n_pad <= kSortTile);
    float4* sbox = reinterpret_cast<float4*>(smem_raw + (size_t)n_pad * 8);        // fast path only: boxes right after the keys
    unsigned* smask = reinterpret_cast<unsigned*>(sbox + m);                        // bitmask path: tiles after the boxes
    const int mask_tile_cap = fast ? (int)(((size_t)a.region_bytes - (size_t)n_pad * 8 - (size_t)m * 16) / 128) : 0;
    // kept positions (16 bit) of the pipelined path, class by class: after the boxes (fast) or, when only the keys fit shared
    // memory ("mid": up to 16384 candidates, boxes gathered from the read-only input by index), right after the keys
    const bool mid = !fast && in_smem && m <= 65535 && ((size_t)n_pad * 8 + (size_t)m * 2 + 64 <= (size_t)a.region_bytes);
    unsigned short* kpos = fast ? reinterpret_cast<unsigned short*>(sbox + m) : reinterpret_cast<unsigned short*>(skeys + n_pad);
    unsigned long long* keys = in_smem ? skeys : (a.ws_keys + (size_t)blockIdx.x * a.n_pad_full);
    __syncthreads();
    if (tid == 0) s_count = 0;
    __syncthreads();

    NMS_STAMP(1);
    // ---- A1 (segment sort): the histogram already gives every class segment's place, so the keys are scattered straight into
    // their segment (counting sort on the label; any order inside a segment) and phase B only sorts WITHIN segments ----------
    const bool seg_sort = in_smem && a.seg_sort;
    if (seg_sort) {
        if (warp == 0) {
            int running = 0;
            for (int base = 0; base <= ncl; base += 32) {
                const int c = base + lane;
                const int v = (c < ncl) ? hist[c_lo + c] : 0;
                int incl = v;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                if (c <= ncl) cls_start[c] = running + incl - v;
                if (c < ncl) cls_count[c] = running + incl - v;               // scatter cursor of the class
                running += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        for (int i = tid; i < nb; i += kNmsThreads) {
            const float s = scores[i]; const int lab = labels[i];
            if (s > a.conf && lab >= c_lo && lab < c_hi) {
                const int pos = atomicAdd(&cls_count[lab - c_lo], 1);
                keys[pos] = ((unsigned long long)(lab - c_lo) << 52) | ((unsigned long long)desc_score_bits(s) << 20) | (unsigned long long)i;
            }
        }
    } else
    // ---- A1: compacted keys (any order: the keys are unique and get sorted) ----------------------
    for (int i0 = 0; i0 < nb; i0 += kNmsThreads) {
        const int i = i0 + tid;
        bool v = false; unsigned long long key = 0;
        if (i < nb) {
            const float s = scores[i]; const int lab = labels[i];
            if (s > a.conf && lab >= c_lo && lab < c_hi) {
                v = true;
                key = ((unsigned long long)(lab - c_lo) << 52) | ((unsigned long long)desc_score_bits(s) << 20) | (unsigned long long)i;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (v) keys[base + __popc(bal & ((1u << lane) - 1u))] = key;
    }
    for (int i = m + tid; i < n_pad; i += kNmsThreads) keys[i] = kInvalidKey;
    __syncthreads();

    NMS_STAMP(2);
    // ---- B: sort -----------------------------------------------------------------------
    if (seg_sort) {
        // segments of more than kWarpSortMax keys: the whole CTA, one after the other; the rest: one warp per segment, all
        // in parallel and without block-wide barriers (bench workload: 54 -> ~15 kcycles for ~2100 keys in ~20 classes)
        // Larger segments: runs of kWarpSortMax keys are sorted by warps in registers like the small segments, then merged
        // pairwise, level by level, by RANK -- every key finds its place in the merged run with one binary search over the
        // sibling run (keys are unique), all keys of all large segments in parallel, ping-pong between the key array and the
        // (still unused) box region.  Measured: a 2962-key class 110 -> ~15 kcycles, the bench workload's 300..750-key classes
        // ~20 -> ~8 kcycles each against the block-wide shared-memory bitonic network this replaces (one barrier per step).
        unsigned long long* scratch = skeys + n_pad;
        const bool can_merge = (size_t)n_pad * 8 + (size_t)m * 8 <= (size_t)a.region_bytes;
        if (!can_merge) {
            for (int c = 0; c < ncl; ++c) {
                const int s0 = cls_start[c], len = cls_start[c + 1] - s0;
                if (len > kWarpSortMax) segment_sort<true>(skeys + s0, len, tid, kNmsThreads);
            }
        }
        for (int c = warp; c < ncl; c += kNmsWarps) {
            const int s0 = cls_start[c], len = cls_start[c + 1] - s0;
            if (len <= kWarpSortMax) warp_sort_segment(skeys + s0, len, lane);
        }
        int maxlen = 0;
        if (can_merge) {
            int item = 0;
            for (int c = 0; c < ncl; ++c) {                    // every warp walks the (short) class table: uniform, no atomics
                const int s0 = cls_start[c], len = cls_start[c + 1] - s0;
                if (len <= kWarpSortMax) continue;
                maxlen = max(maxlen, len);
                for (int r0 = 0; r0 < len; r0 += kWarpSortMax, ++item)
                    if ((item & (kNmsWarps - 1)) == warp) warp_sort_regs<kWarpSortMax / 32>(skeys + s0 + r0, min(kWarpSortMax, len - r0), lane);
            }
        }
        __syncthreads();
        NMS_STAMP(12);
        if (maxlen) {
            bool in_scratch = false;
            for (int w = kWarpSortMax; w < maxlen; w <<= 1) {
                const unsigned long long* src = in_scratch ? scratch : skeys;
                unsigned long long* dst = in_scratch ? skeys : scratch;
                for (int c = 0; c < ncl; ++c) {
                    const int s0 = cls_start[c], len = cls_start[c + 1] - s0;
                    if (len <= kWarpSortMax) continue;
                    for (int e = tid; e < len; e += kNmsThreads) {
                        const int l0 = e & ~(2 * w - 1);                       // first key of the pair of runs
                        const int r0 = min(l0 + w, len), r1 = min(l0 + 2 * w, len);
                        const unsigned long long key = src[s0 + e];
                        const bool left = e < r0;
                        int lo = left ? r0 : l0, hi = left ? r1 : r0;          // sibling run [lo, hi): count its keys below mine
                        const int base = lo;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (src[s0 + mid] < key) lo = mid + 1; else hi = mid; }
                        dst[s0 + l0 + (e - (left ? l0 : r0)) + (lo - base)] = key;
                    }
                }
                __syncthreads();
                in_scratch = !in_scratch;
            }
            if (in_scratch) {
                for (int c = 0; c < ncl; ++c) {
                    const int s0 = cls_start[c], len = cls_start[c + 1] - s0;
                    if (len <= kWarpSortMax) continue;
                    for (int e = tid; e < len; e += kNmsThreads) skeys[s0 + e] = scratch[s0 + e];
                }
                __syncthreads();
            }
        }
    } else if (in_smem) {
        for (int k = 2; k <= n_pad; k <<= 1) bitonic_tile_steps(skeys, n_pad, 0, k, k >> 1);
    } else {
        const int tile = kSortTile, ntiles = n_pad / tile;
        for (int t = 0; t < ntiles; ++t) {       // sort every tile (alternating directions)
            for (int i = tid; i < tile; i += kNmsThreads) skeys[i] = keys[t * tile + i];
            __syncthreads();
            for (int k = 2; k <= tile; k <<= 1) bitonic_tile_steps(skeys, tile, t * tile, k, k >> 1);
            for (int i = tid; i < tile; i += kNmsThreads) keys[t * tile + i] = skeys[i];
            __syncthreads();
        }
        for (int k = tile << 1; k <= n_pad; k <<= 1) {
            for (int j = k >> 1; j >= tile; j >>= 1) {       // wide strides in global memory
                for (int t = tid; t < (n_pad >> 1); t += kNmsThreads) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    int l = i | j;
                    bool up = ((i & k) == 0);
                    unsigned long long x = keys[i], y = keys[l];
                    if ((x > y) == up) { keys[i] = y; keys[l] = x; }
                }
                __syncthreads();
            }
            for (int t = 0; t < ntiles; ++t) {
                for (int i = tid; i < tile; i += kNmsThreads) skeys[i] = keys[t * tile + i];
                __syncthreads();
                bitonic_tile_steps(skeys, tile, t * tile, k, tile >> 1);
                for (int i = tid; i < tile; i += kNmsThreads) keys[t * tile + i] = skeys[i];
                __syncthreads();
            }
        }
    }

    NMS_STAMP(3);
    // ---- C: class segment table (local class index = key >> 52) --------------------------------
    for (int i = tid; i <= m; i += kNmsThreads) {
        int lab_prev = (i == 0) ? -1 : (int)(keys[i - 1] >> 52);
        int lab = (i == m) ? ncl : (int)(keys[i] >> 52);
        for (int c = lab_prev + 1; c <= lab; ++c) cls_start[c] = i;
    }
    int nonfinite = 0;
    if (fast) for (int i = tid; i < m; i += kNmsThreads) {
        const float4 bx = boxes[(int)(keys[i] & kIdxMask)];
        sbox[i] = bx;
        nonfinite |= !(fabsf(bx.x) <= 3.0e38f && fabsf(bx.y) <= 3.0e38f && fabsf(bx.z) <= 3.0e38f && fabsf(bx.w) <= 3.0e38f);
        nonfinite |= !(bx.z >= bx.x && bx.w >= bx.y);                 // inverted boxes: the bitmask path's pre-filter assumes w, h >= 0
    }
    nonfinite = __syncthreads_or(nonfinite);
    NMS_STAMP(4);
    // ---- bitmask path: decision.  Needs every box finite and proper and room for at least two full tile rows of the
    // largest class after the keys and boxes (the masks are built in rounds of as many row blocks as fit) ---------------
    int use_mask = 0;
    if (fast) {
        int wmax = 0;
        for (int c = tid; c < ncl; c += kNmsThreads) wmax = max(wmax, (cls_start[c + 1] - cls_start[c] + 31) >> 5);
        wmax = __reduce_max_sync(0xffffffffu, wmax);
        if (lane == 0 && wmax) atomicMax(&s_last, wmax);
        for (int c = tid; c <= ncl; c += kNmsThreads) state[c] = 0u;                  // next row block of every class
        for (int i = tid; i < (m >> 5) + ncl + 2; i += kNmsThreads) remw_all[i] = 0u;
        __syncthreads();
        wmax = s_last;
        use_mask = (!nonfinite && mask_tile_cap >= 2 * wmax && mask_tile_cap >= 64 && wmax <= a.mask_tile_limit) ? 1 : 0;
        __syncthreads();
        if (tid == 0) s_last = 0;
    } else if (mid) {
        for (int i = tid; i < (m >> 5) + ncl + 2; i += kNmsThreads) remw_all[i] = 0u;   // removed masks of the pipelined path's chunks
    }

    if (use_mask) {
        // ---- D (bitmask), in rounds: plan -> tiles (all warps) -> greedy replay (one warp per class) ------------------
        NMS_STAMP(10);
        for (;;) {
            if (tid == 0) {                                // plan: per class as many further row blocks as the tile budget allows
                int budget = mask_tile_cap, total = 0, done = 1;
                for (int c = 0; c < ncl; ++c) {
                    const int W = (cls_start[c + 1] - cls_start[c] + 31) >> 5;
                    int ke = (int)state[c];
                    chunk_base[c] = total;
                    while (ke < W && W - ke <= budget) { budget -= W - ke; total += W - ke; ++ke; }
                    cls_count[c] = ke;
                    if (ke < W) done = 0;
                }
                chunk_base[ncl] = total;
                s_last = done; s_next = 0;
            }
            __syncthreads();
            const int total_tiles = chunk_base[ncl];
            for (int t = warp; t < total_tiles; t += kNmsWarps) {
                int lo = 0, hi = ncl - 1;                  // largest c with tile_base[c] <= t (classes without tiles share the next base)
                while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (chunk_base[mid] <= t) lo = mid; else hi = mid - 1; }
                const int c = lo, s0 = cls_start[c], nseg = cls_start[c + 1] - s0, W = (nseg + 31) >> 5;
                int u = t - chunk_base[c], k = (int)state[c], rowlen = W - k;
                while (u >= rowlen) { u -= rowlen; --rowlen; ++k; }
                unsigned* out = smask + (size_t)t * 32;
                const unsigned valid = (nseg - 32 * k >= 32) ? 0xffffffffu : ((1u << (nseg - 32 * k)) - 1u);
                if ((remw_all[(s0 >> 5) + c + k] & valid) == valid) { out[lane] = 0u; continue; }   // every row already removed in an earlier round
                if (u == 0) nms_mask_tile<true>(sbox + s0, nseg, k, k, out, a.thr_f, lane);
                else nms_mask_tile<false>(sbox + s0, nseg, k, k + u, out, a.thr_f, lane);
            }
            __syncthreads();
            for (;;) {                                     // big classes first: their replay is the critical path of the round
                int v = 0;
                if (lane == 0) v = atomicAdd(&s_next, 1);
                v = __shfl_sync(0xffffffffu, v, 0);
                if (v >= 2 * ncl) break;
                const int c = (v < ncl) ? v : v - ncl;
                if ((cls_start[c + 1] - cls_start[c] > 128) != (v < ncl)) continue;
                const int kb = (int)state[c], ke = cls_count[c];
                if (ke > kb) {
                    const int s0 = cls_start[c], nseg = cls_start[c + 1] - s0;
                    nms_mask_sweep_rows(smask + (size_t)chunk_base[c] * 32, remw_all + (s0 >> 5) + c, nseg, kb, ke, lane);
                    if (lane == 0) state[c] = (unsigned)ke;
                }
            }
            __syncthreads();
            if (s_last) break;
        }
        NMS_STAMP(11);
        NMS_STAMP(5);
        if (tid == 0) s_next = 0;
        __syncthreads();
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_next, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= ncl) break;
            const int s0 = cls_start[c], nseg = cls_start[c + 1] - s0;
            const int kept = nseg ? nms_mask_compact(skeys + s0, remw_all + (s0 >> 5) + c, nseg, lane) : 0;
            if (lane == 0) cls_count[c] = kept;
        }
    } else {
        // ---- D (pipelined chunks), any size.  fast (m <= 8192): keys and boxes in shared memory.  Otherwise the boxes are
        // re-gathered from the read-only input by index (L1 / L2 hits) and, beyond 16384 candidates, the keys live in the
        // CTA's global workspace (written and read by this CTA only: __threadfence_block orders them like shared memory).
        // A skewed class distribution -- one class holding most of 33 600 anchors -- therefore still runs on all 32 warps
        // (the first version had a one-warp-per-class path there: 104 ms for a 17 000-box class, 35 ms on the 1280x1280
        // MS-Block leg of bench.py) ----
        // Chunks are enumerated class by class in DESCENDING class size (perm, kept in cls_count until the end): the chain of
        // a class's chunks is sequential, so the longest chains must start first and overlap with everything else.
        int* perm = cls_count;
        for (int c = tid; c < ncl; c += kNmsThreads) {
            const int nme = cls_start[c + 1] - cls_start[c];
            int rank = 0;
            for (int o = 0; o < ncl; ++o) {
                const int no = cls_start[o + 1] - cls_start[o];
                rank += (no > nme || (no == nme && o < c)) ? 1 : 0;
            }
            perm[rank] = c;
        }
        __syncthreads();
        if (warp == 0) {                                   // chunk_base[i] = exclusive scan of ceil(n / 32) over perm order
            int running = 0;
            for (int base = 0; base < ncl; base += 32) {
                const int i = base + lane;
                const int c = (i < ncl) ? perm[i] : 0;
                const int v = (i < ncl) ? ((cls_start[c + 1] - cls_start[c] + 31) >> 5) : 0;
                int incl = v;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                if (i < ncl) chunk_base[i] = running + incl - v;
                running += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) chunk_base[ncl] = running;
        }
        for (int c = tid; c < ncl; c += kNmsThreads) state[c] = 0u;
        __syncthreads();
        const int total_chunks = chunk_base[ncl];
        auto run_chunks = [&](auto finite_tag) {
            constexpr bool kFinite = decltype(finite_tag)::value;
            for (int ci = warp; ci < total_chunks; ci += kNmsWarps) {
                int lo = 0, hi = ncl - 1;                      // largest c with chunk_base[c] <= ci
                while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (chunk_base[mid] <= ci) lo = mid; else hi = mid - 1; }
                const int c = perm[lo], j = ci - chunk_base[lo];
                const int s0 = cls_start[c], s1 = cls_start[c + 1];
                const int pos = s0 + 32 * j + lane;
                const bool have = pos < s1;
                const unsigned long long key = have ? keys[pos] : kInvalidKey;
                float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) bj = fast ? sbox[pos] : __ldg(boxes + (int)(key & kIdxMask));
                const float aj = box_area(bj);
                bool removed = !have;
                int applied = 0, kept = 0;
#ifdef YMS_PROF
                long long* tr = (a.prof && blockIdx.x == 0 && lo == 0 && j < 128) ? a.prof + 16 * gridDim.x + 4 * j : nullptr;
                if (tr && lane == 0) tr[0] = clock64();
#endif
                for (;;) {
                    unsigned st = 0;
                    if (lane == 0) st = state[c];
                    st = __shfl_sync(0xffffffffu, st, 0);
                    __threadfence_block();
                    kept = (int)(st & 0xffffu);
                    if (fast) removed = apply_kept_smem<kFinite>(sbox + s0, applied, kept, bj, aj, removed, a.thr_f);
                    else removed = apply_kept([&](int q) { return __ldg(boxes + (int)(keys[s0 + q] & kIdxMask)); }, applied, kept, bj, aj, removed, a.thr_f, lane);
                    applied = kept;
                    const int ahead = j - (int)(st >> 16);     // chunks of this class that are not final yet
                    if (ahead == 0) break;                     // every earlier chunk of this class is final
                    // Back-off by distance: only the chunk whose predecessor is being resolved polls fast.  With a flat 32 ns
                    // sleep the up to 31 waiting warps of the CTA (7 per scheduler, ~18 instructions + a fence per poll)
                    // saturated the issue ports and starved the one warp per class that is on the critical chain.
                    __nanosleep(ahead <= 1 ? 32u : (unsigned)min(a.poll_ns * ahead, 4096));
                }
#ifdef YMS_PROF
                if (tr && lane == 0) { tr[1] = clock64(); tr[3] = kept; }
#endif
                const unsigned surv = resolve_chunk(bj, aj, removed, a.thr_f, lane);
                __syncwarp();                                  // every lane has read its chunk box before the in-place compaction
                if (!removed) {
                    const int dst = s0 + kept + __popc(surv & ((1u << lane) - 1u));   // in place: dst < s0 + 32*(j+1)
                    keys[dst] = key;
                    if (fast) sbox[dst] = bj;
                }
                __threadfence_block();
                __syncwarp();
                if (lane == 0) state[c] = ((unsigned)(j + 1) << 16) | (unsigned)(kept + __popc(surv));
#ifdef YMS_PROF
                if (tr && lane == 0) tr[2] = clock64();
#endif
            }
        };
        // Fast path (keys, boxes and kept positions in shared memory).  The chain of a class is sequential -- chunk j can only be
        // resolved once chunk j-1 is final -- and, measured (profiling build, per-chunk time stamps): one chain step cost 4.6 kcycles
        // on a 93-chunk class (MS-Block leg: 430 of the CTA's 790 kcycles) and 12 kcycles on the bench workload's 12-chunk classes,
        // of which only ~1 kcycle is inherent.  Everything that does not depend on the predecessor's RESULT is therefore done
        // while waiting: H (bit i of lane l: box i of chunk j-1 would suppress my box l -- against every box of the predecessor
        // that is still alive, not just its final survivors) and M (row i: which later boxes of my own chunk box i would
        // suppress).  Once the predecessor publishes its survivor mask S, the step is  removed |= (H & S) != 0;  a 32-step sweep
        // over M on bit masks;  publish.  Every chunk keeps its current removed mask in a word of shared memory (cw[j]; final: ~S),
        // so H only counts rows that can still survive; both masks are built right after the chunk's first pass over the kept list.  Sorted boxes and keys stay
        // immutable (the kept list holds 16-bit POSITIONS), so a chunk may read its predecessor's boxes at any time.
        auto run_chunks_fast = [&](auto finite_tag, auto smem_tag) {
            constexpr bool kFinite = decltype(finite_tag)::value;
            constexpr bool kBoxSmem = decltype(smem_tag)::value;
            auto bx = [&](int q) -> float4 { return kBoxSmem ? sbox[q] : __ldg(boxes + (int)(keys[q] & kIdxMask)); };
            for (int ci = warp; ci < total_chunks; ci += kNmsWarps) {
                int lo = 0, hi = ncl - 1;                      // largest c with chunk_base[c] <= ci
                while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (chunk_base[mid] <= ci) lo = mid; else hi = mid - 1; }
                const int c = perm[lo], j = ci - chunk_base[lo];
                const int s0 = cls_start[c], s1 = cls_start[c + 1];
                const int pos = s0 + 32 * j + lane;
                const bool have = pos < s1;
                float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) bj = bx(pos);
                const float aj = box_area(bj);
                bool removed = !have;
                int applied = 0, kept = 0;
                unsigned H = 0u, Mrow = 0u, s_pred = 0u;
                bool pre = false;                              // H and M have been computed
                volatile unsigned* cw = remw_all + (s0 >> 5) + c;             // removed masks of the class's chunks (zeroed above)
                unsigned published = 0u;
#ifdef YMS_PROF
                long long* tr = (a.prof && blockIdx.x == 0 && lo == 0 && j < 128) ? a.prof + 16 * gridDim.x + 4 * j : nullptr;
                if (tr && lane == 0) tr[0] = clock64();
#endif
                for (;;) {
                    unsigned st = 0;
                    if (lane == 0) st = state[c];
                    st = __shfl_sync(0xffffffffu, st, 0);
                    __threadfence_block();
                    kept = (int)(st & 0xffffu);
                    const int ahead = j - (int)(st >> 16);     // chunks of this class that are not final yet
                    int lim = kept;
                    if (ahead == 0 && pre) {                   // the last popc(S) kept boxes are the predecessor's survivors: H covers them
                        s_pred = ~cw[j - 1];
                        lim = kept - __popc(s_pred);
                    }
                    removed = apply_kept_pos<kFinite>(bx, kpos + s0, applied, lim, bj, aj, removed, a.thr_f);
                    applied = lim;
                    if (ahead == 0) break;
                    const unsigned rem_now = __ballot_sync(0xffffffffu, removed);
                    if (rem_now != published) { published = rem_now; if (lane == 0) cw[j] = rem_now; }
                    if (!pre) {
                        // right after the first pass over the kept list: only the boxes of this chunk that are still alive need
                        // rows / columns.  H is built TRANSPOSED -- lane i holds box i of the predecessor, the loop runs over my
                        // alive boxes l, the ballot over the predecessor's lanes IS H of lane l -- so both masks cost one pair
                        // test per alive box of this chunk, not per box of the predecessor.
                        pre = true;
                        const float4 bp = bx(s0 + 32 * (j - 1) + lane);        // the predecessor is a full chunk
                        const float ap = box_area(bp);
                        const bool p_alive = !((cw[j - 1] >> lane) & 1u);      // predecessor's boxes that may still survive (superset of S)
                        const unsigned alive0 = ~rem_now;
                        unsigned todo = alive0;
                        while (todo) {                                         // warp-uniform
                            const int i = __ffs(todo) - 1;
                            todo &= todo - 1u;
                            float4 bi;                                         // my box i, broadcast
                            bi.x = __shfl_sync(0xffffffffu, bj.x, i); bi.y = __shfl_sync(0xffffffffu, bj.y, i);
                            bi.z = __shfl_sync(0xffffffffu, bj.z, i); bi.w = __shfl_sync(0xffffffffu, bj.w, i);
                            const float ai = box_area(bi);
                            const unsigned hcol = __ballot_sync(0xffffffffu, p_alive && pair_hit<kFinite>(bp, ap, bi, ai, a.thr_f));
                            const unsigned row = __ballot_sync(0xffffffffu, (lane > i) && pair_hit<kFinite>(bi, ai, bj, aj, a.thr_f));
                            if (lane == i) { H = hcol; Mrow = row; }
                        }
                        continue;                                              // the chain has moved on meanwhile: poll again at once
                    }
                    // Back-off by distance; the chunk next in line spins (one warp per class: cheap, and its wake-up latency is
                    // chain latency)
                    if (ahead > 1) __nanosleep((unsigned)min(a.poll_ns * ahead, 4096));
                }
#ifdef YMS_PROF
                if (tr && lane == 0) { tr[1] = clock64(); tr[3] = kept; }
#endif
                unsigned alive;
                if (pre) {
                    removed = removed || ((H & s_pred) != 0u);
                    alive = ~__ballot_sync(0xffffffffu, removed);
                    #pragma unroll
                    for (int b8 = 0; b8 < 4; ++b8) {
                        if (((alive >> (8 * b8)) & 0xffu) == 0u) continue;     // warp-uniform
                        unsigned r[8];
                        #pragma unroll
                        for (int q = 0; q < 8; ++q) r[q] = __shfl_sync(0xffffffffu, Mrow, 8 * b8 + q);
                        #pragma unroll
                        for (int q = 0; q < 8; ++q) if ((alive >> (8 * b8 + q)) & 1u) alive &= ~r[q];
                    }
                } else {
                    alive = resolve_chunk(bj, aj, removed, a.thr_f, lane);
                }
                if ((alive >> lane) & 1u) kpos[s0 + kept + __popc(alive & ((1u << lane) - 1u))] = (unsigned short)pos;
                if (lane == 0) cw[j] = ~alive;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) state[c] = ((unsigned)(j + 1) << 16) | (unsigned)(kept + __popc(alive));
#ifdef YMS_PROF
                if (tr && lane == 0) tr[2] = clock64();
#endif
            }
        };
        if (fast) { if (nonfinite) run_chunks_fast(std::false_type{}, std::true_type{}); else run_chunks_fast(std::true_type{}, std::true_type{}); }
        else if (mid) run_chunks_fast(std::false_type{}, std::false_type{});
        else run_chunks(std::false_type{});
        __syncthreads();
        for (int c = tid; c < ncl; c += kNmsThreads) cls_count[c] = (int)(state[c] & 0xffffu);
    }
    __syncthreads();

    NMS_STAMP(6);
    // ---- E: per-CTA keep list, then the last CTA of the image concatenates -------------------------
    if (warp == 0) {
        int running = 0;
        for (int base = 0; base < ncl; base += 32) {
            const int c = base + lane;
            const int v = (c < ncl) ? cls_count[c] : 0;
            int incl = v;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (c < ncl) cls_count[c] = running + incl - v;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) cls_count[ncl] = running;
    }
    __syncthreads();
    const int my_total = cls_count[ncl];
    int32_t* out = (G == 1) ? (a.keep + (size_t)b * n) : (a.ws_stage + (size_t)blockIdx.x * n);
    for (int c = warp; c < ncl; c += kNmsWarps) {
        const int off = cls_count[c], cnt = cls_count[c + 1] - off, s0 = cls_start[c];
        if ((fast && !use_mask) || mid) for (int r = lane; r < cnt; r += 32) out[off + r] = (int32_t)(keys[kpos[s0 + r]] & kIdxMask);
        else for (int r = lane; r < cnt; r += 32) out[off + r] = (int32_t)(keys[s0 + r] & kIdxMask);
    }
    if (G == 1) {
        for (int i = my_total + tid; i < n; i += kNmsThreads) out[i] = -1;
        if (tid == 0) a.keep_count[b] = my_total;
        return;
    }
    NMS_STAMP(7);
    if (tid == 0) a.ws_count[blockIdx.x] = my_total;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned ticket = atomicAdd(&a.ws_ticket[b], 1u);
        s_last = (ticket == (unsigned)(G - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int32_t* keep = a.keep + (size_t)b * n;
    int off = 0;
    for (int q = 0; q < G; ++q) {
        const int cnt = __ldcg(a.ws_count + b * G + q);
        const int32_t* src = a.ws_stage + (size_t)(b * G + q) * n;
        for (int r = tid; r < cnt; r += kNmsThreads) keep[off + r] = __ldcg(src + r);
        off += cnt;
    }
    for (int i = off + tid; i < n; i += kNmsThreads) keep[i] = -1;
    if (tid == 0) a.keep_count[b] = off;
    NMS_STAMP(8);
}

__global__ void gather_dets_kernel(const float4* boxes, const float* scores, const int32_t* labels,
                                   const int32_t* keep, const int32_t* keep_count, int n, int max_det,
                                   float* dets) {
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= max_det) return;
    float* o = dets + ((size_t)b * max_det + r) * 6;
    const int cnt = min(keep_count[b], max_det);
    if (r < cnt) {
        int i = keep[(size_t)b * n + r];
        float4 bx = boxes[(size_t)b * n + i];
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        o[4] = scores[(size_t)b * n + i];
        o[5] = (float)labels[(size_t)b * n + i];
    } else {
        o[0] = o[1] = o[2] = o[3] = o[4] = 0.f; o[5] = -1.f;
    }
}

int next_pow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

int pick_groups(int batch, int num_classes, int n) {
    const int forced = g_opt.nms_groups;
    int g = forced > 0 ? forced : kNumSMs / batch;
    const int by_size = (n + 4095) / 4096;          // keep the expected candidates per CTA well inside the shared-memory fast path
    if (forced <= 0 && g < by_size) g = by_size;
    if (g < 1) g = 1;
    if (g > kMaxGroups) g = kMaxGroups;
    if (g > num_classes) g = num_classes;
    return g;
}

struct WsLayout { size_t keys, stage, count, ticket, total; };
WsLayout ws_layout(int batch, int n, int groups) {
    WsLayout w;
    const int n_pad = next_pow2(n);
    size_t off = 0;
    w.keys = off;   off += (n_pad > kSortTile) ? (size_t)batch * groups * n_pad * 8 : 0;
    w.stage = off;  off += (groups > 1) ? (((size_t)batch * groups * n * 4 + 15) & ~(size_t)15) : 0;
    w.count = off;  off += ((size_t)batch * groups * 4 + 15) & ~(size_t)15;
    w.ticket = off; off += ((size_t)batch * 4 + 15) & ~(size_t)15;
    w.total = off;
    return w;
}

}  // namespace

}  // namespace yms

using namespace yms;

extern "C" size_t yms_nms_workspace_bytes(int batch, int n) {
    if (batch <= 0 || n <= 0) return 0;
    return ws_layout(batch, n, pick_groups(batch, kMaxGroups, n)).total;   // group count before the num_classes clamp (upper bound)
}

extern "C" int yms_nms_batched(const float* boxes, const float* scores, const int32_t* labels,
                               const int32_t* n_valid, int batch, int n, int num_classes,
                               float conf_thr, double iou_thr, int32_t* keep, int32_t* keep_count,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (batch < 0 || n < 0 || num_classes <= 0) return fail(YMS_E_ARG, "nms: bad sizes");
    if (batch == 0) return 0;
    if (n > (1 << 20) || num_classes > kMaxClasses)
        return fail(YMS_E_UNSUPPORTED, "nms: N <= 2^20 and num_classes <= %d required", kMaxClasses);
    if (!keep_count || (n > 0 && (!boxes || !scores || !labels || !keep))) return fail(YMS_E_ARG, "nms: null pointer");
    if (!(iou_thr >= 0.0)) return fail(YMS_E_ARG, "nms: iou_thr must be >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * batch, st);
        return e == cudaSuccess ? 0 : fail((int)e, "nms: memset failed");
    }
    if (((uintptr_t)boxes & 15) != 0) return fail(YMS_E_ARG, "nms: boxes must be 16-byte aligned");
    const int groups = pick_groups(batch, num_classes, n);
    const WsLayout w = ws_layout(batch, n, groups);
    if (w.total > 0 && (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)))
        return fail(YMS_E_WORKSPACE, "nms: workspace %zu < %zu (or misaligned)", workspace_bytes, w.total);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    NmsArgs a;
    a.boxes = reinterpret_cast<const float4*>(boxes); a.scores = scores; a.labels = labels; a.n_valid = n_valid;
    a.n = n; a.num_classes = num_classes; a.groups = groups; a.conf = conf_thr;
    // largest float <= iou_thr: (double)ovr > thr  <=>  ovr > thr_f for every float ovr
    float tf = (float)iou_thr;
    if ((double)tf > iou_thr) tf = nextafterf(tf, -INFINITY);
    a.thr_f = tf;
    a.keep = keep; a.keep_count = keep_count;
    a.n_pad_full = next_pow2(n);
    a.ws_keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
    a.ws_stage = reinterpret_cast<int32_t*>(ws + w.stage);
    a.ws_count = reinterpret_cast<int32_t*>(ws + w.count);
    a.ws_ticket = reinterpret_cast<unsigned int*>(ws + w.ticket);
    a.prof = g_prof_buf;
    a.mask_tile_limit = g_opt.nms_mask_tiles;
    a.poll_ns = g_opt.nms_poll_ns < 16 ? 16 : g_opt.nms_poll_ns;
    a.seg_sort = g_opt.nms_sort_bitonic ? 0 : 1;
    if (groups > 1) {
        cudaError_t e = cudaMemsetAsync(a.ws_ticket, 0, sizeof(unsigned int) * batch, st);
        if (e != cudaSuccess) return fail((int)e, "nms: ticket memset failed");
    }
    const size_t tables = (size_t)(num_classes + 2) * 4 * 4 + (size_t)(kSortTile / 32 + num_classes + 4) * 4 + 16;   // class tables + removed words
    if (kKeyRegionBytes + tables > 232448) return fail(YMS_E_UNSUPPORTED, "nms: %d classes need %zu bytes of shared memory (limit 232448)", num_classes, kKeyRegionBytes + tables);
    // keys | boxes | kept positions: 192 KB hold the sort tile and 8192 candidates without the position list; what the class tables
    // leave of another 16 KB lets a full 8192-candidate CTA stay on the fast path
    size_t region = (232448 - tables) & ~(size_t)15;
    if (region > kKeyRegionBytes + (size_t)kFastCap * 2 + 64) region = kKeyRegionBytes + (size_t)kFastCap * 2 + 64;
    a.region_bytes = (int)region;
    const size_t smem = region + tables;
    static std::atomic<size_t> attr_bytes[64];           // per device: cudaFuncSetAttribute is a per-device setting
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (smem > attr_bytes[dev_id & 63].load()) {
        cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail((int)e, "nms: smem attribute: %s", cudaGetErrorString(e));
        attr_bytes[dev_id & 63].store(smem);
    }
    nms_kernel<<<batch * groups, kNmsThreads, smem, st>>>(a);
    return check_launch("nms_kernel");
}

extern "C" int yms_gather_detections(const float* boxes, const float* scores, const int32_t* labels,
                                     const int32_t* keep, const int32_t* keep_count, int batch, int n,
                                     int max_det, float* dets, void* stream) {
    if (batch <= 0 || max_det <= 0) return 0;
    dim3 grid(ceil_div(max_det, 128), batch);
    gather_dets_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(boxes), scores, labels,
                                                                keep, keep_count, n, max_det, dets);
    return check_launch("gather_dets_kernel");
}
