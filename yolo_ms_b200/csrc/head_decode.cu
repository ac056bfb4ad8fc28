// Head decode: DFL softmax-expectation + anchor grid + stride scaling + class sigmoid, plus the
// per-anchor candidate pre-filter inputs (xyxy, best score, best class), in ONE kernel.
//
// Replaces the eval branch of Head.forward (yolov8/model/yolov8_head.py:127-144),
// Head.make_anchors (:146-158), DFL.forward (yolov8/model/components.py:176-191) and the
// candidate selection of the post-process (yolov8/tools/test.py:166-179) of the reference.
// All arithmetic is fp32 (bf16 has 4 px resolution at 640 px -- SURVEY.md section 4).
//
// HBM-bound: per anchor it reads (64+nc) logits and writes 4+nc floats.  A CTA stages a tile of
// 32 anchors x (64+nc) logits in shared memory with 16-byte coalesced loads, computes from
// shared memory, stages the [32, 4+nc] output tile and streams it out with 16-byte stores.
#include "decode_math.cuh"

#include <stdlib.h>

namespace yms {
namespace {

constexpr int kTileAnchors = 32;
constexpr int kDecodeThreads = 256;

struct DecodeArgs {
    const void* raw[3];
    int hw[3];            // anchors per scale
    int w[3];             // width per scale
    float stride[3];
    int tiles_before[4];  // prefix of tiles per image over scales
    int anchor_base[3];
    int batch, nc, no, total_anchors, tiles_per_image;
    float* pred; float4* cand_boxes; float* cand_scores; int32_t* cand_labels;
};

// sigmoid_f, to_xyxy, dfl_expectation, dfl_box, cls_chunk16: decode_math.cuh (shared with the fused conv epilogue)

template <typename T> struct RawLoad;
template <> struct RawLoad<float> { static constexpr int kVec = 4; };           // elements per 16 B
template <> struct RawLoad<__nv_bfloat16> { static constexpr int kVec = 8; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(kDecodeThreads) head_decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int no = a.no, nc = a.nc, nout = 4 + nc;
    T* s_in = reinterpret_cast<T*>(smem_raw);                                    // [32][no]
    float* s_out = reinterpret_cast<float*>(smem_raw + ((kTileAnchors * no * sizeof(T) + 15) & ~15));  // [32][nout]
    float* s_dist = s_out + kTileAnchors * nout;                                 // [32][4]

    const int img = blockIdx.x / a.tiles_per_image;
    const int tile = blockIdx.x % a.tiles_per_image;
    int sc = 0;
    if (tile >= a.tiles_before[1]) sc = 1;
    if (tile >= a.tiles_before[2]) sc = 2;
    const int a0 = (tile - a.tiles_before[sc]) * kTileAnchors;      // first anchor of the tile inside the scale
    const int cnt = min(kTileAnchors, a.hw[sc] - a0);
    const T* src = reinterpret_cast<const T*>(a.raw[sc]) + ((size_t)img * a.hw[sc] + a0) * no;
    const int tid = threadIdx.x;

    // ---- stage the logits (contiguous cnt*no elements) ----
    {
        constexpr int V = RawLoad<T>::kVec;
        const int nelem = cnt * no;
        if ((((uintptr_t)src) & 15) == 0) {
            const int nvec = nelem / V;
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_in);
            for (int i = tid; i < nvec; i += kDecodeThreads) d4[i] = __ldg(s4 + i);
            for (int i = nvec * V + tid; i < nelem; i += kDecodeThreads) s_in[i] = src[i];
        } else {
            for (int i = tid; i < nelem; i += kDecodeThreads) s_in[i] = src[i];
        }
    }
    __syncthreads();

    // ---- DFL: one thread per (anchor, side) ----
    if (tid < cnt * 4) {
        const int an = tid >> 2, side = tid & 3;
        const T* l = s_in + an * no + side * kRegMax;
        float v[kRegMax];
        float mx = -INFINITY;
        #pragma unroll
        for (int k = 0; k < kRegMax; ++k) { v[k] = to_f(l[k]); mx = fmaxf(mx, v[k]); }
        float sum = 0.f, wsum = 0.f;
        #pragma unroll
        for (int k = 0; k < kRegMax; ++k) { const float e = __expf(v[k] - mx); sum += e; wsum = fmaf((float)k, e, wsum); }
        s_dist[tid] = wsum / sum;                  // sum_k k * softmax_k
    }
    // ---- class sigmoid ----
    for (int i = tid; i < cnt * nc; i += kDecodeThreads) {
        const int an = i / nc, c = i - an * nc;
        s_out[an * nout + 4 + c] = sigmoid_f(to_f(s_in[an * no + 4 * kRegMax + c]));
    }
    __syncthreads();

    // ---- boxes (yolov8_head.py:139-143) ----
    if (tid < cnt) {
        const int g = a0 + tid;
        const float ax = (float)(g % a.w[sc]) + 0.5f, ay = (float)(g / a.w[sc]) + 0.5f;
        const float st = a.stride[sc];
        const float x1 = ax - s_dist[tid * 4 + 0], y1 = ay - s_dist[tid * 4 + 1];
        const float x2 = ax + s_dist[tid * 4 + 2], y2 = ay + s_dist[tid * 4 + 3];
        float* o = s_out + tid * nout;
        o[0] = ((x1 + x2) / 2.0f) * st;
        o[1] = ((y1 + y2) / 2.0f) * st;
        o[2] = (x2 - x1) * st;
        o[3] = (y2 - y1) * st;
    }
    __syncthreads();

    const size_t out_anchor = (size_t)img * a.total_anchors + a.anchor_base[sc] + a0;
    // ---- candidates: one warp per anchor (first max wins, like torch.max) ----
    if (a.cand_boxes) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int an = warp; an < cnt; an += kDecodeThreads / 32) {
            const float* o = s_out + an * nout;
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int c = lane; c < nc; c += 32) {
                float v = o[4 + c];
                if (v > best || (bi == 0x7fffffff)) { best = v; bi = c; }
            }
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                a.cand_boxes[out_anchor + an] = to_xyxy(o[0], o[1], o[2], o[3]);
                a.cand_scores[out_anchor + an] = best;
                a.cand_labels[out_anchor + an] = bi;
            }
        }
    }
    // ---- stream the output tile (contiguous cnt*nout floats, 16-byte aligned since nout%4==0) ----
    float* dst = a.pred + out_anchor * nout;
    const int nelem = cnt * nout;
    if ((nout & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(s_out);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < (nelem >> 2); i += kDecodeThreads) d4[i] = s4[i];
    } else {
        for (int i = tid; i < nelem; i += kDecodeThreads) dst[i] = s_out[i];
    }
}

// ---- v2: one thread per (anchor, 16-channel group), logits straight from global memory ----------------
// ncu on the kernel above: 108 M warp instructions for 270 MB (issue-bound at 34 % of HBM): scalar shared
// memory staging, an integer division per class score and warp-shuffle arg-max dominated.  Here a tile of
// 32 anchors x (64 + nc) logits is 32 x G chunks of 16 channels (G = 4 DFL sides + nc/16 class groups); thread
// (anchor, g) loads ITS 16 logits with 16-byte loads (consecutive threads read consecutive 64 B: coalesced),
// keeps them in registers, and produces either one DFL distance or 16 sigmoid scores + their local arg-max.
// One thread per anchor then assembles the box and the candidate; the [32, 4+nc] output tile is staged in
// shared memory and streamed out with 16-byte stores.  ~10x fewer instructions per anchor.
template <typename T>
__device__ __forceinline__ void load16(const T* src, float (&v)[16]);
template <>
__device__ __forceinline__ void load16<float>(const float* src, float (&v)[16]) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    #pragma unroll
    for (int q = 0; q < 4; ++q) { const float4 t = __ldg(s4 + q); v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
}
template <>
__device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* src, float (&v)[16]) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint4 t = __ldg(s4 + q);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
        #pragma unroll
        for (int j = 0; j < 4; ++j) { v[8 * q + 2 * j] = bf16_lo(w[j]); v[8 * q + 2 * j + 1] = bf16_hi(w[j]); }
    }
}

constexpr int kMaxGroupsV2 = 4 + 16;            // nc <= 256

template <typename T>
__global__ void __launch_bounds__(kTileAnchors * kMaxGroupsV2) head_decode_v2_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int no = a.no, nc = a.nc, nout = 4 + nc, ngroups = no >> 4, ncg = ngroups - 4;
    float* s_out = reinterpret_cast<float*>(smem_raw);                           // [32][nout]
    float* s_dist = s_out + kTileAnchors * nout;                                 // [32][4]
    float* s_best = s_dist + kTileAnchors * 4;                                   // [32][ncg]
    int* s_bidx = reinterpret_cast<int*>(s_best + kTileAnchors * ncg);           // [32][ncg]

    const int img = blockIdx.x / a.tiles_per_image;
    const int tile = blockIdx.x % a.tiles_per_image;
    int sc = 0;
    if (tile >= a.tiles_before[1]) sc = 1;
    if (tile >= a.tiles_before[2]) sc = 2;
    const int a0 = (tile - a.tiles_before[sc]) * kTileAnchors;
    const int cnt = min(kTileAnchors, a.hw[sc] - a0);
    const T* src = reinterpret_cast<const T*>(a.raw[sc]) + ((size_t)img * a.hw[sc] + a0) * no;
    const int tid = threadIdx.x;
    const int an = tid / ngroups, g = tid - an * ngroups;

    if (an < cnt) {
        float v[16];
        load16<T>(src + (size_t)tid * 16, v);                                    // chunk tid of the contiguous tile
        if (g < 4) {                                                             // DFL side g: sum_k k * softmax_k
            s_dist[an * 4 + g] = dfl_expectation(v);
        } else {                                                                 // 16 class scores
            const int c0 = (g - 4) * 16;
            float best; int bi;
            float4 r[4];
            cls_chunk16(v, c0, r, best, bi);                                     // first max wins (torch.max)
            float4* o4 = reinterpret_cast<float4*>(s_out + an * nout + 4 + c0);
            #pragma unroll
            for (int q = 0; q < 4; ++q) o4[q] = r[q];
            s_best[an * ncg + (g - 4)] = best;
            s_bidx[an * ncg + (g - 4)] = bi;
        }
    }
    __syncthreads();

    const size_t out_anchor = (size_t)img * a.total_anchors + a.anchor_base[sc] + a0;
    if (tid < cnt) {                                                             // boxes (yolov8_head.py:139-143) + candidate
        const int gidx = a0 + tid;
        const float ax = (float)(gidx % a.w[sc]) + 0.5f, ay = (float)(gidx / a.w[sc]) + 0.5f;
        const float st = a.stride[sc];
        const float4 d = *reinterpret_cast<const float4*>(s_dist + tid * 4);
        const float4 box = dfl_box(ax, ay, d, st);
        *reinterpret_cast<float4*>(s_out + tid * nout) = box;
        if (a.cand_boxes) {
            float best = s_best[tid * ncg]; int bi = s_bidx[tid * ncg];
            for (int q = 1; q < ncg; ++q) {
                const float ob = s_best[tid * ncg + q];
                if (ob > best) { best = ob; bi = s_bidx[tid * ncg + q]; }
            }
            a.cand_boxes[out_anchor + tid] = to_xyxy(box.x, box.y, box.z, box.w);
            a.cand_scores[out_anchor + tid] = best;
            a.cand_labels[out_anchor + tid] = bi;
        }
    }
    __syncthreads();
    const float4* s4 = reinterpret_cast<const float4*>(s_out);
    float4* d4 = reinterpret_cast<float4*>(a.pred + out_anchor * nout);
    const int nvec = (cnt * nout) >> 2;
    for (int i = tid; i < nvec; i += blockDim.x) d4[i] = s4[i];
}

// pred [B*A, 4+nc] -> candidates; one warp per anchor, rows are read coalesced.
__global__ void __launch_bounds__(256) select_candidates_kernel(const float* pred, long long rows, int nc,
                                                                 float4* boxes, float* scores, int32_t* labels) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int nout = 4 + nc;
    for (long long r = warp; r < rows; r += nwarps) {
        const float* o = pred + r * nout;
        float best = -INFINITY; int bi = 0x7fffffff;
        for (int c = lane; c < nc; c += 32) {
            float v = __ldg(o + 4 + c);
            if (v > best || (bi == 0x7fffffff)) { best = v; bi = c; }
        }
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, off);
            int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
        }
        if (lane == 0) {
            boxes[r] = to_xyxy(__ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(o + 3));
            scores[r] = best;
            labels[r] = bi;
        }
    }
}

}  // namespace
}  // namespace yms

using namespace yms;

extern "C" int yms_head_decode(const void* raw0, const void* raw1, const void* raw2, int raw_dtype,
                               int batch, const int32_t* host_hw, int num_classes,
                               const float* host_strides, float* pred, float* cand_boxes,
                               float* cand_scores, int32_t* cand_labels, void* stream) {
    if (batch < 0 || num_classes <= 0 || !host_hw || !host_strides) return fail(YMS_E_ARG, "decode: bad arguments");
    if (batch == 0) return 0;
    if (!raw0 || !raw1 || !raw2 || !pred) return fail(YMS_E_ARG, "decode: null pointer");
    if (cand_boxes && (!cand_scores || !cand_labels)) return fail(YMS_E_ARG, "decode: candidate outputs must come together");
    if (raw_dtype != YMS_DTYPE_F32 && raw_dtype != YMS_DTYPE_BF16) return fail(YMS_E_UNSUPPORTED, "decode: raw dtype");
    DecodeArgs a;
    a.raw[0] = raw0; a.raw[1] = raw1; a.raw[2] = raw2;
    a.batch = batch; a.nc = num_classes; a.no = 4 * kRegMax + num_classes;
    int tiles = 0, anchors = 0;
    for (int i = 0; i < 3; ++i) {
        int h = host_hw[2 * i], w = host_hw[2 * i + 1];
        if (h <= 0 || w <= 0) return fail(YMS_E_ARG, "decode: bad feature size");
        a.hw[i] = h * w; a.w[i] = w; a.stride[i] = host_strides[i];
        a.tiles_before[i] = tiles; a.anchor_base[i] = anchors;
        tiles += ceil_div(h * w, kTileAnchors); anchors += h * w;
    }
    a.tiles_before[3] = tiles; a.tiles_per_image = tiles; a.total_anchors = anchors;
    a.pred = pred; a.cand_boxes = reinterpret_cast<float4*>(cand_boxes);
    a.cand_scores = cand_scores; a.cand_labels = cand_labels;
    const size_t esz = raw_dtype == YMS_DTYPE_F32 ? 4 : 2;
    size_t smem = ((kTileAnchors * a.no * esz + 15) & ~(size_t)15) + (size_t)kTileAnchors * (4 + num_classes) * 4 + kTileAnchors * 4 * 4;
    if (smem > 200 * 1024) return fail(YMS_E_UNSUPPORTED, "decode: num_classes too large");
    const long long grid = (long long)batch * tiles;
    if (grid > 0x7fffffffLL) return fail(YMS_E_UNSUPPORTED, "decode: grid too large");
    cudaError_t e;
    if ((num_classes % 16) == 0 && num_classes <= 256) {      // v2: one thread per 16-channel chunk
        const int ngroups = a.no / 16, ncg = ngroups - 4;
        const size_t smem2 = (size_t)kTileAnchors * ((4 + num_classes) + 4 + 2 * ncg) * 4;
        const int threads = kTileAnchors * ngroups;
        if (raw_dtype == YMS_DTYPE_F32) head_decode_v2_kernel<float><<<(unsigned)grid, threads, smem2, (cudaStream_t)stream>>>(a);
        else head_decode_v2_kernel<__nv_bfloat16><<<(unsigned)grid, threads, smem2, (cudaStream_t)stream>>>(a);
        return check_launch("head_decode_v2_kernel");
    }
    if (raw_dtype == YMS_DTYPE_F32) {
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(head_decode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return fail((int)e, "decode: smem attribute");
        head_decode_kernel<float><<<(unsigned)grid, kDecodeThreads, smem, (cudaStream_t)stream>>>(a);
    } else {
        if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(head_decode_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return fail((int)e, "decode: smem attribute");
        head_decode_kernel<__nv_bfloat16><<<(unsigned)grid, kDecodeThreads, smem, (cudaStream_t)stream>>>(a);
    }
    return check_launch("head_decode_kernel");
}

extern "C" int yms_select_candidates(const float* pred, int batch, int anchors, int num_classes,
                                     float* boxes, float* scores, int32_t* labels, void* stream) {
    if (batch < 0 || anchors < 0 || num_classes <= 0) return fail(YMS_E_ARG, "select: bad sizes");
    const long long rows = (long long)batch * anchors;
    if (rows == 0) return 0;
    if (!pred || !boxes || !scores || !labels) return fail(YMS_E_ARG, "select: null pointer");
    if (((uintptr_t)boxes & 15) != 0) return fail(YMS_E_ARG, "select: boxes must be 16-byte aligned");
    long long blocks = (rows + 7) / 8;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    select_candidates_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, rows, num_classes,
        reinterpret_cast<float4*>(boxes), scores, labels);
    return check_launch("select_candidates_kernel");
}
