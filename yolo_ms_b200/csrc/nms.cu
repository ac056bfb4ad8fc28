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
//   B  bitonic sort of the keys in shared memory (size = next pow2 of ITS candidate count)
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

namespace yms {
namespace {

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kSortTile = 16384;           // u64 keys that fit the 128 KB key region
constexpr int kFastCap = 8192;             // fast path: 64 KB keys + 128 KB boxes
constexpr int kMaxClasses = 2047;
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

struct NmsArgs {
    const float4* boxes; const float* scores; const int32_t* labels; const int32_t* n_valid;
    int n, num_classes, groups; float conf; float thr_f;
    int32_t* keep; int32_t* keep_count;
    unsigned long long* ws_keys;   // [B*G][pow2(n)] (only when pow2(n) > kSortTile)
    int32_t* ws_stage;             // [B*G][n]  per-CTA keep lists
    int32_t* ws_count;             // [B*G]
    unsigned int* ws_ticket;       // [B] zeroed by the host before the launch
    int n_pad_full;                // pow2(n)
};

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(NmsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);
    float4* sbox = reinterpret_cast<float4*>(smem_raw + (size_t)kFastCap * 8);      // fast path only
    const int G = a.groups;
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int c_lo = (int)((long long)g * a.num_classes / G), c_hi = (int)((long long)(g + 1) * a.num_classes / G);
    const int ncl = c_hi - c_lo;                                                    // classes owned by this CTA
    int* cls_start = reinterpret_cast<int*>(smem_raw + kKeyRegionBytes);            // [ncl + 1]
    int* cls_count = cls_start + (ncl + 1);                                         // [ncl + 1] kept counts, then offsets
    int* chunk_base = cls_count + (ncl + 1);                                        // [ncl + 1]
    volatile unsigned* state = reinterpret_cast<volatile unsigned*>(chunk_base + (ncl + 1));   // [ncl] (done chunks<<16 | kept)
    __shared__ int s_count, s_next, s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.n;
    const int nb = a.n_valid ? min(max(a.n_valid[b], 0), n) : n;
    const float4* boxes = a.boxes + (size_t)b * n;
    const float* scores = a.scores + (size_t)b * n;
    const int32_t* labels = a.labels + (size_t)b * n;

    if (tid == 0) { s_count = 0; s_next = 0; s_last = 0; }
    __syncthreads();

    // ---- A0: count this CTA's candidates --------------------------------------------------
    {
        int cnt = 0;
        for (int i = tid; i < nb; i += kNmsThreads) {
            const float s = scores[i]; const int lab = labels[i];
            cnt += (s > a.conf && lab >= c_lo && lab < c_hi) ? 1 : 0;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&s_count, cnt);
    }
    __syncthreads();
    const int m = s_count;
    int n_pad = 32;
    while (n_pad < m) n_pad <<= 1;
    const bool fast = (m <= kFastCap);
    const bool in_smem = (n_pad <= kSortTile);
    unsigned long long* keys = in_smem ? skeys : (a.ws_keys + (size_t)blockIdx.x * a.n_pad_full);
    __syncthreads();
    if (tid == 0) s_count = 0;
    __syncthreads();

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

    // ---- B: sort -----------------------------------------------------------------------
    if (in_smem) {
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

    // ---- C: class segment table (local class index = key >> 52) --------------------------------
    for (int i = tid; i <= m; i += kNmsThreads) {
        int lab_prev = (i == 0) ? -1 : (int)(keys[i - 1] >> 52);
        int lab = (i == m) ? ncl : (int)(keys[i] >> 52);
        for (int c = lab_prev + 1; c <= lab; ++c) cls_start[c] = i;
    }
    if (fast) for (int i = tid; i < m; i += kNmsThreads) sbox[i] = boxes[(int)(keys[i] & kIdxMask)];
    __syncthreads();

    if (fast) {
        // ---- D (fast): pipelined chunks -------------------------------------------------------
        if (warp == 0) {                                   // chunk_base = exclusive scan of ceil(n_c / 32)
            int running = 0;
            for (int base = 0; base < ncl; base += 32) {
                const int c = base + lane;
                const int v = (c < ncl) ? ((cls_start[c + 1] - cls_start[c] + 31) >> 5) : 0;
                int incl = v;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                if (c < ncl) chunk_base[c] = running + incl - v;
                running += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) chunk_base[ncl] = running;
        }
        for (int c = tid; c < ncl; c += kNmsThreads) state[c] = 0u;
        __syncthreads();
        const int total_chunks = chunk_base[ncl];
        for (int ci = warp; ci < total_chunks; ci += kNmsWarps) {
            int lo = 0, hi = ncl - 1;                      // largest c with chunk_base[c] <= ci
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (chunk_base[mid] <= ci) lo = mid; else hi = mid - 1; }
            const int c = lo, j = ci - chunk_base[c];
            const int s0 = cls_start[c], s1 = cls_start[c + 1];
            const int pos = s0 + 32 * j + lane;
            const bool have = pos < s1;
            const unsigned long long key = have ? skeys[pos] : kInvalidKey;
            const float4 bj = have ? sbox[pos] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float aj = box_area(bj);
            bool removed = !have;
            int applied = 0, kept = 0;
            for (;;) {
                unsigned st = 0;
                if (lane == 0) st = state[c];
                st = __shfl_sync(0xffffffffu, st, 0);
                __threadfence_block();
                kept = (int)(st & 0xffffu);
                removed = apply_kept([&](int q) { return sbox[s0 + q]; }, applied, kept, bj, aj, removed, a.thr_f, lane);
                applied = kept;
                if ((int)(st >> 16) == j) break;           // every earlier chunk of this class is final
                if (applied == kept) __nanosleep(64);
            }
            const unsigned surv = resolve_chunk(bj, aj, removed, a.thr_f, lane);
            if (!removed) {
                const int dst = s0 + kept + __popc(surv & ((1u << lane) - 1u));   // in place: dst < s0 + 32*(j+1)
                skeys[dst] = key;
                sbox[dst] = bj;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) state[c] = ((unsigned)(j + 1) << 16) | (unsigned)(kept + __popc(surv));
        }
        __syncthreads();
        for (int c = tid; c < ncl; c += kNmsThreads) cls_count[c] = (int)(state[c] & 0xffffu);
    } else {
        // ---- D (general): one warp per class segment, boxes gathered from global memory -----------
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_next, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= ncl) break;
            const int s0 = cls_start[c], s1 = cls_start[c + 1];
            int kept = 0;
            for (int d = s0; d < s1; d += 32) {
                const int pos = d + lane;
                const bool have = pos < s1;
                const unsigned long long key = have ? keys[pos] : kInvalidKey;
                float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) bj = boxes[(int)(key & kIdxMask)];
                const float aj = box_area(bj);
                bool removed = !have;
                removed = apply_kept([&](int q) { return boxes[(int)(keys[s0 + q] & kIdxMask)]; }, 0, kept, bj, aj, removed, a.thr_f, lane);
                const unsigned surv = resolve_chunk(bj, aj, removed, a.thr_f, lane);
                if (!removed) keys[s0 + kept + __popc(surv & ((1u << lane) - 1u))] = key;
                kept += __popc(surv);
                __syncwarp();
            }
            if (lane == 0) cls_count[c] = kept;
        }
    }
    __syncthreads();

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
        for (int r = lane; r < cnt; r += 32) out[off + r] = (int32_t)(keys[s0 + r] & kIdxMask);
    }
    if (G == 1) {
        for (int i = my_total + tid; i < n; i += kNmsThreads) out[i] = -1;
        if (tid == 0) a.keep_count[b] = my_total;
        return;
    }
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

int pick_groups(int batch, int num_classes) {
    int g = kNumSMs / batch;
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
    return ws_layout(batch, n, pick_groups(batch, kMaxGroups)).total;   // group count before the num_classes clamp (upper bound)
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
    const int groups = pick_groups(batch, num_classes);
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
    if (groups > 1) {
        cudaError_t e = cudaMemsetAsync(a.ws_ticket, 0, sizeof(unsigned int) * batch, st);
        if (e != cudaSuccess) return fail((int)e, "nms: ticket memset failed");
    }
    const int ncl_max = (num_classes + groups - 1) / groups + 1;
    const size_t smem = kKeyRegionBytes + (size_t)(ncl_max + 1) * 4 * 4 + 16;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(kKeyRegionBytes + (size_t)(kMaxClasses + 2) * 16 + 16));
        if (e != cudaSuccess) return fail((int)e, "nms: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
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
