// Shared host-side plan structure and tensor-map encoding for the convolution kernels.
#pragma once
#include "conv_epilogue.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <utility>

struct yms_conv_plan;

namespace yms {

constexpr int kBlockK = 64;                    // bf16 channels per k-block = one 128 B swizzle row

// Head decode fused into the epilogue of a head branch's final 1x1 convolution (conv_gemm.cu, yms_conv_plan_fuse_decode):
// the accumulator row of a thread is one anchor, so DFL / sigmoid / arg-max are thread-local.
struct DecodeFuse {
    int mode;                                  // 0 off, 1 box branch (4 x 16 DFL logits), 2 class branch (nc logits)
    int hw, w;                                 // anchors (pixels) and width of this scale's map
    int anchor_base, anchors;                  // first anchor of the scale, anchors per image over all scales
    int nout;                                  // 4 + nc floats per prediction row
    const float* stride;                       // DEVICE f32: this scale's stride, read at run time (head.stride is honoured per call)
    float* pred; float4* cand_boxes; float* cand_scores; int* cand_labels;
};

struct ConvKernelParams {
    int tiles_x, tiles_y, batch;               // M tiling (output space)
    int tw, th;                                // output pixels per tile (tw*th <= 128)
    int out_w, out_h;
    int n_tiles, block_n, c_out;
    int kb1, kb2, c_in1, c_in2;                // 64-channel blocks per tap of source 1 / 2
    int taps, ksize, stride;
    int act, out_f32, has_res;
    int num_stages, total_tiles;
    int resident;                              // all weight tiles of the (single) N tile stay in smem
    int s2_dense;                              // stride-2 input is a dense NHWC tensor: 5-D parity view, no traversal stride
    int pair;                                  // CTA-pair kernel (cta_group::2, M = 256): two consecutive M tiles per cluster, each CTA half of B
    int m_tiles;                               //   M tiles (tiles_x * tiles_y * batch); an odd count leaves the last cluster's second CTA without a tile
    int acc_stages;                            // TMEM accumulator stages == active epilogue groups (1, 2 or 4)
    int epi_groups;                            // epilogue groups of 4 warps: 4 (one 576-thread CTA per SM) or 2 ("half" CTAs: 320
                                               // threads, <= 113 KB of shared memory, 256 TMEM columns, so that TWO CTAs -- of this
                                               // layer or of the next one, whose prologue then overlaps this layer's tail -- share an SM)
    int tmem_cols;                             // 128 * epi_groups
    uint32_t mg_n_tiles, mg_tiles_x, mg_tiles_y;  // fast_div magics
    int bias_pad;                              // floats of shared-memory bias (c_out rounded up to 64)
    const float* bias;
    float* y_f32; long long y_ps;
    const float* up; long long up_ps;          // fp32 partial sums at HALF resolution added before bias / activation (or null),
    int up_w, up_hw;                           //   pixel stride in floats; W and H*W of THIS conv's (full-resolution) map
    DecodeFuse dec;                            // dec.mode != 0: the epilogue writes predictions / candidates instead of y
    long long* prof;                           // YMS_PROF builds: [grid][16] cycle counters (else unused)
};


// ---- 3x3 stride-1 halo kernel (conv3x3.cu) ----
struct Conv3Params {
    int tiles_x, tiles_y, batch;               // sub-tile grid (8 x th output pixels each)
    int th;                                    // rows per sub-tile (<= 16)
    int sub;                                   // sub-tiles (adjacent in x) per work item: 1 or 2
    int super_x;                               // ceil(tiles_x / sub)
    int out_w, out_h;
    int n_tiles, block_n, c_out, c_in, kb;     // kb = 64-channel blocks of c_in
    int act, has_res;
    int a_stages, b_stages, resident;          // resident: all weights of the (single) N tile stay in smem
    int acc_stages;                            // TMEM accumulator stages (2 if sub*block_n <= 256)
    int total_items;
    int bias_pad;
    int desc_mode;                             // 0: base_offset = 0, 1: base_offset = (addr >> 7) & 7
    int s2pair;                                // stride-2 pair-line mode (c_in == 32, dense input): see conv3x3.cu
    int pair;                                  // CTA-pair kernel (cta_group::2, M = 256): two x-adjacent sub-tiles per cluster, each CTA half of B
    int vy, vh, bands;                         // pair kernel, virtual-row tiling: images vh = H + 2 virtual rows apart, `bands` bands of 16 rows
    uint32_t mg_vh;
    int epi_groups;                            // pair kernel: epilogue groups of 4 warps (4, or 2 when that is what lets the weights stay resident)
    int planes;                                // TMA boxes per A stage (== sub, or 2 parity planes in s2pair mode)
    int wtiles;                                // resident weight tiles per 64-channel block (9 taps, or 6 pair-packed tiles)
    int pitch;                                 // lines per halo row (10, or 9 pixel pairs in s2pair mode)
    uint32_t mg_n_tiles, mg_super_x, mg_tiles_y;  // fast_div magics
    uint32_t halo_bytes;                       // 10 * (th + 2) * 128
    int halo_stage;                            // smem stride between halo tiles (== halo_bytes, multiple of 128)
    const float* bias;
    long long* prof;                           // YMS_PROF builds: [grid][16] cycle counters (else unused)
};

extern long long* g_prof_buf;                  // set by yms_debug_set_prof (capi.cu)

inline uint32_t fast_div_magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d); }

// Launch with the programmatic-stream-serialization attribute (kernels call pdl_wait() before touching
// global data).  The library option pdl_off disables it.  `cluster` > 1: thread-block clusters of that many CTAs along x.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, int cluster, Args&&... args) {
    const bool enabled = !g_opt.pdl_off;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (enabled) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_pdl_cluster(kernel, grid, block, smem, stream, 1, std::forward<Args>(args)...);
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode();
int encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* addr, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr, const char* what);
int encode_act(CUtensorMap* m, const void* ptr, int c, int64_t ps, int batch, int h, int w, bool flat,
               int box_x, int box_y, int estride, const char* what, bool f32 = false);

int conv3_plan_init(::yms_conv_plan* pl, const yms_conv_params* q);   // conv3x3.cu
int conv3_s2pair_plan_init(::yms_conv_plan* pl, const yms_conv_params* q);
int conv3_plan_run(const ::yms_conv_plan* pl, cudaStream_t stream);

}  // namespace yms

struct yms_conv_plan {
    CUtensorMap tm_x, tm_x2, tm_w, tm_y, tm_res, tm_res2, tm_y2;
    int kind;                   // 0: generic implicit GEMM, 1: 3x3 stride-1 halo kernel
    yms::ConvKernelParams kp;
    yms::Conv3Params k3;
    int grid, threads;
    size_t smem;
    double flops, bytes;
};

