// Batched class-aware NMS for the YOLO post-process, one kernel, one CTA per image.
//
// Replaces yolov8/tools/test.py:181-218 (confidence filter + `for c in unique(cls): nms(...)`
// around torchvision.ops.nms) of the reference.  Results are bit-identical to that path:
//   * candidates: score > conf (strict, fp32)
//   * order: label ascending, score descending, ties by lower index (== stable sort per class)
//   * suppression iff (double)IoU > iou_thr, IoU = inter / (area_i + area_j - inter) in fp32
//     with every operation individually rounded (no FMA contraction), NaN never suppresses.
//
// Phases inside the kernel (all in shared memory for N <= 16384 candidates per image):
//   A  build 64-bit sort keys  [removed:1 | label:11 | ~score:32 | index:20]
//   B  bitonic sort of the keys (smem tiles; global passes only when N > 16384)
//   C  per-class segment table from the sorted keys
//   D  greedy suppression, one warp per class segment, target-chunk-major:
//      a chunk of 32 sorted boxes lives in the lanes; previously kept boxes of the segment
//      are streamed past it (shfl broadcast), then the chunk is resolved in-register;
//      __ballot_sync/__popc compact the survivors in place at the segment front
//   E  exclusive scan of per-class survivor counts, write keep indices in output order
#include "common.cuh"

namespace yms {
namespace {

constexpr int kNmsThreads = 1024;
constexpr int kSortTile = 16384;           // u64 keys held in shared memory (128 KB)
constexpr int kMaxClasses = 2047;
constexpr unsigned long long kInvalidKey = ~0ull;
constexpr unsigned long long kIdxMask = (1ull << 20) - 1;

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

struct NmsArgs {
    const float4* boxes; const float* scores; const int32_t* labels; const int32_t* n_valid;
    int n, n_pad, num_classes; float conf; float thr_f;
    int32_t* keep; int32_t* keep_count; unsigned long long* ws_keys;   // ws: [B][n_pad] if n_pad > tile
};

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(NmsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tile = min(a.n_pad, kSortTile);
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(smem_raw);
    int* cls_start = reinterpret_cast<int*>(skeys + tile);       // [num_classes + 1]
    int* cls_count = cls_start + (a.num_classes + 1);            // [num_classes + 1] (then offsets)
    __shared__ int s_valid, s_next_class;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int n = a.n;
    const int nb = a.n_valid ? min(max(a.n_valid[b], 0), n) : n;
    const float4* boxes = a.boxes + (size_t)b * n;
    const float* scores = a.scores + (size_t)b * n;
    const int32_t* labels = a.labels + (size_t)b * n;
    const bool in_smem = (a.n_pad <= kSortTile);
    unsigned long long* keys = in_smem ? skeys : (a.ws_keys + (size_t)b * a.n_pad);

    if (tid == 0) { s_valid = 0; s_next_class = 0; }
    __syncthreads();

    // ---- A: keys ---------------------------------------------------------------------
    int my_valid = 0;
    for (int i = tid; i < a.n_pad; i += blockDim.x) {
        unsigned long long key = kInvalidKey;
        if (i < nb) {
            float s = scores[i];
            int lab = labels[i];
            if (s > a.conf && lab >= 0 && lab < a.num_classes) {
                key = ((unsigned long long)lab << 52) | ((unsigned long long)desc_score_bits(s) << 20) |
                      (unsigned long long)i;
                ++my_valid;
            }
        }
        keys[i] = key;
    }
    my_valid = __reduce_add_sync(0xffffffffu, my_valid);
    if (lane == 0 && my_valid) atomicAdd(&s_valid, my_valid);
    __syncthreads();
    const int m = s_valid;                       // number of candidates

    // ---- B: sort -----------------------------------------------------------------------
    if (in_smem) {
        for (int k = 2; k <= a.n_pad; k <<= 1) bitonic_tile_steps(skeys, tile, 0, k, k >> 1);
    } else {
        const int ntiles = a.n_pad / tile;
        for (int t = 0; t < ntiles; ++t) {       // sort every tile (alternating directions)
            for (int i = tid; i < tile; i += blockDim.x) skeys[i] = keys[t * tile + i];
            __syncthreads();
            for (int k = 2; k <= tile; k <<= 1) bitonic_tile_steps(skeys, tile, t * tile, k, k >> 1);
            for (int i = tid; i < tile; i += blockDim.x) keys[t * tile + i] = skeys[i];
            __syncthreads();
        }
        for (int k = tile << 1; k <= a.n_pad; k <<= 1) {
            for (int j = k >> 1; j >= tile; j >>= 1) {       // wide strides in global memory
                for (int t = tid; t < (a.n_pad >> 1); t += blockDim.x) {
                    int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    int l = i | j;
                    bool up = ((i & k) == 0);
                    unsigned long long x = keys[i], y = keys[l];
                    if ((x > y) == up) { keys[i] = y; keys[l] = x; }
                }
                __syncthreads();
            }
            for (int t = 0; t < ntiles; ++t) {
                for (int i = tid; i < tile; i += blockDim.x) skeys[i] = keys[t * tile + i];
                __syncthreads();
                bitonic_tile_steps(skeys, tile, t * tile, k, tile >> 1);
                for (int i = tid; i < tile; i += blockDim.x) keys[t * tile + i] = skeys[i];
                __syncthreads();
            }
        }
    }

    // ---- C: class segment table ----------------------------------------------------------
    for (int i = tid; i <= m; i += blockDim.x) {
        int lab_prev = (i == 0) ? -1 : (int)(keys[i - 1] >> 52);
        int lab = (i == m) ? a.num_classes : (int)(keys[i] >> 52);
        for (int c = lab_prev + 1; c <= lab; ++c) cls_start[c] = i;
    }
    __syncthreads();

    // ---- D: greedy suppression, one warp per class segment ------------------------------------
    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&s_next_class, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= a.num_classes) break;
        const int s0 = cls_start[c], s1 = cls_start[c + 1];
        int kept = 0;                            // survivors so far, compacted at keys[s0 ...]
        for (int d = s0; d < s1; d += 32) {
            const int pos = d + lane;
            const bool have = pos < s1;
            unsigned long long key = have ? keys[pos] : kInvalidKey;
            float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have) bj = boxes[(int)(key & kIdxMask)];
            const float aj = box_area(bj);
            bool removed = !have;
            // previously kept boxes of this class vs this chunk
            for (int g = 0; g < kept; g += 32) {
                if (__all_sync(0xffffffffu, removed)) break;
                const int cnt = min(32, kept - g);
                float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < cnt) bp = boxes[(int)(keys[s0 + g + lane] & kIdxMask)];
                const float ap = box_area(bp);
                for (int p = 0; p < cnt; ++p) {
                    float4 bi;
                    bi.x = __shfl_sync(0xffffffffu, bp.x, p);
                    bi.y = __shfl_sync(0xffffffffu, bp.y, p);
                    bi.z = __shfl_sync(0xffffffffu, bp.z, p);
                    bi.w = __shfl_sync(0xffffffffu, bp.w, p);
                    float ai = __shfl_sync(0xffffffffu, ap, p);
                    if (!removed && suppresses(bi, ai, bj, aj, a.thr_f)) removed = true;
                }
            }
            // resolve the chunk itself in score order
            unsigned alive = ~__ballot_sync(0xffffffffu, removed);
            for (int i = 0; i < 32; ++i) {
                if (!((alive >> i) & 1u)) continue;            // warp-uniform
                float4 bi;
                bi.x = __shfl_sync(0xffffffffu, bj.x, i);
                bi.y = __shfl_sync(0xffffffffu, bj.y, i);
                bi.z = __shfl_sync(0xffffffffu, bj.z, i);
                bi.w = __shfl_sync(0xffffffffu, bj.w, i);
                float ai = __shfl_sync(0xffffffffu, aj, i);
                bool hit = (lane > i) && !removed && suppresses(bi, ai, bj, aj, a.thr_f);
                if (hit) removed = true;
                alive &= ~__ballot_sync(0xffffffffu, hit);
            }
            // compact survivors in place (positions written are < d + 32: already consumed)
            const unsigned surv = alive;
            if (!removed) keys[s0 + kept + __popc(surv & ((1u << lane) - 1u))] = key;
            kept += __popc(surv);
            __syncwarp();
        }
        if (lane == 0) cls_count[c] = kept;
    }
    __syncthreads();

    // ---- E: output offsets and keep list -----------------------------------------------------
    if (tid < 32) {
        int running = 0;
        for (int base = 0; base < a.num_classes; base += 32) {
            int c = base + lane;
            int v = (c < a.num_classes) ? cls_count[c] : 0;
            int incl = v;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (c < a.num_classes) cls_count[c] = running + incl - v;    // exclusive offset
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { cls_count[a.num_classes] = running; a.keep_count[b] = running; }
    }
    __syncthreads();
    int32_t* keep = a.keep + (size_t)b * n;
    const int total = cls_count[a.num_classes];
    const int warp = tid >> 5, nwarps = blockDim.x >> 5;
    for (int c = warp; c < a.num_classes; c += nwarps) {
        const int off = cls_count[c];
        const int cnt = cls_count[c + 1] - off;
        const int s0 = cls_start[c];
        for (int r = lane; r < cnt; r += 32) keep[off + r] = (int32_t)(keys[s0 + r] & kIdxMask);
    }
    for (int i = total + tid; i < n; i += blockDim.x) keep[i] = -1;
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

int next_pow2(int v) { int p = 2; while (p < v) p <<= 1; return p; }

}  // namespace

}  // namespace yms

using namespace yms;

extern "C" size_t yms_nms_workspace_bytes(int batch, int n) {
    if (batch <= 0 || n <= 0) return 0;
    int n_pad = next_pow2(n);
    if (n_pad <= kSortTile) return 0;
    return (size_t)batch * n_pad * sizeof(unsigned long long);
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
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(keep_count, 0, sizeof(int32_t) * batch, (cudaStream_t)stream);
        return e == cudaSuccess ? 0 : fail((int)e, "nms: memset failed");
    }
    if (((uintptr_t)boxes & 15) != 0) return fail(YMS_E_ARG, "nms: boxes must be 16-byte aligned");
    NmsArgs a;
    a.boxes = reinterpret_cast<const float4*>(boxes); a.scores = scores; a.labels = labels; a.n_valid = n_valid;
    a.n = n; a.n_pad = next_pow2(n); a.num_classes = num_classes; a.conf = conf_thr;
    // largest float <= iou_thr: (double)ovr > thr  <=>  ovr > thr_f for every float ovr
    float tf = (float)iou_thr;
    if ((double)tf > iou_thr) tf = nextafterf(tf, -INFINITY);
    a.thr_f = tf;
    a.keep = keep; a.keep_count = keep_count; a.ws_keys = nullptr;
    if (a.n_pad > kSortTile) {
        size_t need = yms_nms_workspace_bytes(batch, n);
        if (!workspace || workspace_bytes < need) return fail(YMS_E_WORKSPACE, "nms: workspace %zu < %zu", workspace_bytes, need);
        a.ws_keys = reinterpret_cast<unsigned long long*>(workspace);
    }
    const int tile = a.n_pad < kSortTile ? a.n_pad : kSortTile;
    size_t smem = (size_t)tile * 8 + (size_t)(num_classes + 1) * 8;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortTile * 8 + (kMaxClasses + 1) * 8);
        if (e != cudaSuccess) return fail((int)e, "nms: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    nms_kernel<<<batch, kNmsThreads, smem, (cudaStream_t)stream>>>(a);
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
