/* yms_b200.h -- C ABI of libyms_b200.so: the B200 (sm_100a) kernels of the YOLO-MS / YOLOv8
 * inference hot path (forward -> DFL decode -> class-aware NMS).
 *
 * Boundary rules:
 *   - extern "C", plain pointers and sizes, no torch types; every pointer is DEVICE memory
 *     unless the parameter name starts with host_.
 *   - stream-ordered and re-entrant per stream: no allocation, no host synchronisation;
 *     `stream` is a cudaStream_t passed as void*.
 *   - every entry point returns 0 on success, a negative YMS_E_* code for argument errors or
 *     a positive cudaError_t; yms_last_error() returns a thread-local message.
 *   - there is NO CPU implementation behind any of these symbols.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository rafaelghiorzi/YOLO-MS).
 */
#ifndef YMS_B200_H_
#define YMS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YMS_ABI_VERSION 1

#define YMS_E_ARG        (-1)   /* invalid argument                                   */
#define YMS_E_UNSUPPORTED (-2)  /* shape / dtype outside what the kernels implement    */
#define YMS_E_WORKSPACE  (-3)   /* workspace too small                                 */
#define YMS_E_DRIVER     (-4)   /* CUDA driver entry point (tensor map encode) failed  */

#define YMS_DTYPE_F32  0
#define YMS_DTYPE_BF16 1

int         yms_abi_version(void);
const char* yms_last_error(void);
/* Number of kernel launches issued through this library by the calling process. */
long long   yms_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Convolution unit: y = act(conv(x, W) + b) (+ residual), NHWC bf16 activations.
 * Replaces Conv.forward (yolov8/model/components.py:69-77: Conv2d(bias=False) -> BatchNorm2d
 * (eval, eps 1e-3) -> SiLU) with BN folded into W/b by the host, the in-place residual of
 * Bottleneck.forward (components.py:87-93) and the biased 1x1 nn.Conv2d that ends each head
 * branch (yolov8/model/yolov8_head.py:86,101).  The tensors may be channel slices of wider
 * buffers (pixel stride != channels), which is how torch.cat / slicing in C2f.forward
 * (components.py:108-122) and Neck.forward (yolov8_neck.py:76-92) become pointer arithmetic.
 * Implementation: tcgen05.mma (TMEM accumulators) implicit GEMM, operands staged by TMA.
 * ------------------------------------------------------------------------------------------ */
typedef struct yms_conv_params {
    /* problem */
    int32_t batch, in_h, in_w;        /* input spatial size                                       */
    int32_t c_in, c_out;              /* channels of the (slice of the) input / output            */
    int32_t ksize;                    /* 1 or 3 (padding = ksize/2)                                */
    int32_t stride;                   /* 1 or 2                                                   */
    int32_t act;                      /* 0 = identity, 1 = SiLU                                   */
    int32_t out_dtype;                /* YMS_DTYPE_BF16 or YMS_DTYPE_F32                           */
    /* second input source (K-concatenation: conv over cat[x, x2], or conv(x + x2) when
       weights are repeated); c_in2 = 0 disables it.  Same spatial size as x.               */
    int32_t c_in2;
    int32_t variant;                  /* kernel variant for 3x3/s1 layers: 0 = library heuristic, 1 = generic implicit GEMM,
                                         2 = halo kernel (1 sub-tile per work item), 3 = halo kernel (2 sub-tiles per work item);
                                         4 = stride-2 pair-line kernel for 3x3/s2 with c_in == 32 on a dense input: `weight` is then
                                         PAIR-PACKED bf16 [6][c_out][64]: tile 2*ky = [w(ky,1) | w(ky,2)], tile 2*ky+1 = [0 | w(ky,0)].
                                         5 = CTA-pair halo kernel (cta_group::2, M = 256: two x-adjacent sub-tiles per 2-CTA cluster, each CTA
                                         half of every weight tile; c_out <= 256, maps at least 9 pixels wide); for every other bf16-output
                                         convolution (1x1, 3x3/s2, two sources) 5 selects the CTA-pair variant of the generic kernel (two
                                         consecutive M tiles per cluster); 6 = that generic CTA-pair kernel for a 3x3/s1 layer.
                                         7 = the CTA-pair halo kernel with virtual-row tiling, for maps whose height is not a multiple
                                         of 16 (40, 20): the images are stacked H + 2 rows apart and cut into bands of exactly 16 rows, a
                                         cluster takes one 8-pixel column of two consecutive bands (c_out <= 256, H >= 16).
                                         Results agree to fp32 accumulation order; used by the host-side per-layer autotuner. */
    /* tensors */
    const void* x;   int64_t x_pixel_stride;      /* bf16, elements between consecutive pixels   */
    const void* x2;  int64_t x2_pixel_stride;
    void*       y;   int64_t y_pixel_stride;      /* bf16 or f32                                 */
    const void* residual; int64_t res_pixel_stride; /* bf16 [.., c_out] added after act, or NULL  */
    const void* weight;   /* bf16 [ksize*ksize][c_out][c_in + c_in2]  (tap-major, K contiguous)   */
    const float* bias;    /* f32 [c_out]                                                          */
} yms_conv_params;

typedef struct yms_conv_plan yms_conv_plan;   /* opaque: encoded tensor maps + launch geometry */

int yms_conv_plan_create(const yms_conv_params* p, yms_conv_plan** plan);
int yms_conv_plan_run(const yms_conv_plan* plan, void* stream);
int yms_conv_plan_destroy(yms_conv_plan* plan);
/* 2*MACs of the plan and the algorithmic bytes (inputs + outputs + weights) it moves. */
int yms_conv_plan_cost(const yms_conv_plan* plan, double* flops, double* bytes);

/* Head decode fused into a head branch's final convolution.  Turns the plan of the biased 1x1 nn.Conv2d that ends
 * head.box[i] / head.cls[i] (yolov8/model/yolov8_head.py:86,101; plan created with ksize 1, act 0, f32 output) into one that
 * does NOT store its logits but decodes them in the epilogue -- DFL.forward (components.py:176-191), the eval branch of
 * Head.forward incl. make_anchors (yolov8_head.py:127-158) and the candidate selection of the post-process
 * (tools/test.py:166-179) -- writing the same values yms_head_decode would produce from the stored logits (bit-identical):
 *   branch 1 (box,   c_out == 64):          pred[b, a, 0:4] = (cx, cy, w, h) * stride,  cand_boxes[b, a] = xyxy
 *   branch 2 (class, c_out == num_classes): pred[b, a, 4:]  = sigmoid(logits),          cand_scores / cand_labels[b, a] = max / first arg-max
 * with a = anchor_base + y * map_w + x.  `stride` points to ONE float in DEVICE memory that is read when the plan runs, so
 * head.stride (a plain attribute in the reference, :79) is honoured at call time even inside a captured CUDA graph.
 * Limits: num_classes a multiple of 16, <= 128.  Returns YMS_E_UNSUPPORTED otherwise (use yms_head_decode then). */
typedef struct yms_decode_fusion {
    int32_t branch;                   /* 1 = box branch, 2 = class branch                                  */
    int32_t map_h, map_w;             /* feature map of this scale == the plan's in_h, in_w                */
    int32_t anchor_base;              /* index of the scale's first anchor among an image's anchors        */
    int32_t anchors;                  /* A: anchors per image over all scales                              */
    int32_t num_classes;
    const float* stride;              /* DEVICE f32[1]                                                     */
    float* pred;                      /* f32 [B, A, 4 + num_classes]                                       */
    float* cand_boxes;                /* f32 [B, A, 4] (branch 1) or NULL                                  */
    float* cand_scores;               /* f32 [B, A]    (branch 2) or NULL                                  */
    int32_t* cand_labels;             /* i32 [B, A]    (branch 2) or NULL; comes with cand_scores          */
} yms_decode_fusion;
int yms_conv_plan_fuse_decode(yms_conv_plan* plan, const yms_decode_fusion* f);

/* Nearest-x2 upsample + channel concat fused into the 1x1 convolution that consumes them (Upsample.forward components.py:159-160,
 * torch.cat in Neck.forward yolov8_neck.py:77-83, C2f.conv1 components.py:108).  A 1x1 convolution is linear per pixel, so
 *     conv1x1(cat[upsample2x(a), b]) = upsample2x(W_a . a) + W_b . b          (before bias / activation)
 * The host runs the first term at HALF resolution as a linear 1x1 plan with f32 output and zero bias (4x fewer MACs, the upsampled
 * tensor never exists); this call makes `plan` (the 1x1 convolution over b alone, bf16 output, at H x W = out_h x out_w) add
 * t[n, y/2, x/2, :] to its accumulator before bias and activation.  t: DEVICE f32 [B, out_h/2, out_w/2, c_out] (pixel stride in
 * floats).  Differs from the unfused computation only in fp32 summation order.  c_out must be a multiple of 16 (of 64 beyond 256
 * channels, where the kernel tiles N); YMS_E_UNSUPPORTED otherwise (run yms_upsample2x + the plain convolution then). */
int yms_conv_plan_add_upsampled(yms_conv_plan* plan, const float* t, int64_t t_pixel_stride, int out_h, int out_w);

/* Stem: first layer backbone.conv0 (yolov8/model/yolov8_backbone.py:39, 3x3 stride 2 on the
 * NCHW fp32 image, components.py:69-77) -> NHWC bf16.  weight f32 [c_out][3][3][3] with BN
 * folded, bias f32 [c_out]. */
int yms_stem_conv(const float* x_nchw, int batch, int in_h, int in_w, int c_out,
                  const float* weight, const float* bias,
                  void* y_nhwc_bf16, int64_t y_pixel_stride, void* stream);

/* Same layer fed by the raw image: x uint8 NHWC [B,H,W,3] (what PIL / cv2 decode), with the reference's
 * pre-processing ToTensor + Normalize(mean, std) (yolov8/tools/test.py:114-119; resize excluded) fused
 * into the gather: v -> (v/255 - mean[c]) / std[c].  SURVEY.md section 8(f) rank 1. */
int yms_stem_conv_u8(const uint8_t* x_nhwc, int batch, int in_h, int in_w, int c_out,
                     const float* weight, const float* bias, const float* host_mean /* 3 */,
                     const float* host_std /* 3 */, void* y_nhwc_bf16, int64_t y_pixel_stride, void* stream);

/* One pass of the resize step of the reference's pre-processing: T.Resize((h, w)) on a PIL image
 * (yolov8/tools/test.py:114-119,142-145) = Pillow's Image.resize(BILINEAR), src/libImaging/Resample.c: a separable
 * two-pass (horizontal, then vertical) convolution with 22-bit fixed-point coefficients, each pass rounding to uint8.
 * src/dst uint8 HWC; bounds int32 [out][2] = (first source index, count), coeffs int32 [out][ksize] (device memory,
 * computed by the host exactly as precompute_coeffs/normalize_coeffs_8bpc do).  horizontal != 0: x pass (dst_h == src_h),
 * else y pass (dst_w == src_w).  SURVEY.md section 8(f) rank 1; bit-exact vs Pillow. */
int yms_resample_u8(const uint8_t* src, int src_h, int src_w, int channels, int64_t src_row_stride_bytes,
                    uint8_t* dst, int dst_h, int dst_w, int64_t dst_row_stride_bytes,
                    const int32_t* bounds, const int32_t* coeffs, int ksize, int horizontal, void* stream);

/* Depthwise k x k (k in 3,5,7,9; stride 1; pad k/2) + folded BN + SiLU, NHWC bf16.
 * Conv(c, c, k, 1, k//2, groups=c) of components.py:69-77 as used by the repo-local MS-Block.
 * weight f32 [k*k][c] (tap-major), bias f32 [c]. */
int yms_dwconv(const void* x, int64_t x_pixel_stride, int batch, int h, int w, int channels, int ksize,
               const float* weight, const float* bias, void* y, int64_t y_pixel_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * MS-Block branch layer (repo-local block built from the reference's Conv unit, components.py:69-77, incl. its `groups`
 * argument; the reference only sketches the block, annotations.md:66-133):
 *        x [c_in (+ c_in2)] --1x1 (pw1)--> e [e_ch] --depthwise k x k--> d [e_ch] --1x1 (pw2)--> y [c_out]
 * every step = Conv2d(bias=False) + BatchNorm2d(eval) folded into weight / bias by the host + SiLU.  One kernel, three modes:
 *   mode 0: depthwise only            y = SiLU(dw(e) + dw_bias)                        (the kernel behind yms_dwconv)
 *   mode 1: depthwise -> pw2          y = act2(W2 . SiLU(dw(e) + dw_bias) + bias2)     (d never reaches HBM)
 *   mode 2: pw1 -> depthwise -> pw2   e = SiLU(W1 . cat[x, x2] + bias1) computed on the halo of every 16 x 8 output tile,
 *                                     zero outside the image (the depthwise padding) -- neither e nor d reaches HBM.
 * Activations NHWC bf16 (channel slices allowed: pixel strides in elements); depthwise in fp32 on the CUDA cores
 * (fma.rn.f32x2), 1x1 convolutions on tcgen05 with fp32 accumulation.
 * Limits: ksize in {3,5,7,9}; channels % 8 == 0; modes 1/2: c_out % 16 == 0, <= 256; mode 2 additionally needs the layer's
 * operands to fit in shared memory / TMEM (small c): YMS_E_UNSUPPORTED otherwise -- fall back to a lower mode.
 * ------------------------------------------------------------------------------------------ */
typedef struct yms_ms_params {
    int32_t mode;                     /* 0, 1, 2 (above)                                                   */
    int32_t batch, h, w;
    int32_t ksize;                    /* depthwise kernel size, padding ksize/2, stride 1                  */
    int32_t e_ch;                     /* expanded (depthwise) channels                                     */
    int32_t c_out;                    /* pw2 output channels (modes 1, 2)                                  */
    int32_t act2;                     /* pw2 activation: 0 identity, 1 SiLU                                */
    int32_t c_in, c_in2;              /* pw1 input channels of x / x2 (mode 2; c_in2 = 0: one source)      */
    const void* e;  int64_t e_pixel_stride;    /* bf16 [B,H,W,e_ch]  (modes 0, 1)                          */
    const void* x;  int64_t x_pixel_stride;    /* bf16 [B,H,W,c_in]  (mode 2)                              */
    const void* x2; int64_t x2_pixel_stride;   /* bf16 [B,H,W,c_in2] (mode 2) or NULL                      */
    void* y;        int64_t y_pixel_stride;    /* bf16 [B,H,W,c_out] (mode 0: [B,H,W,e_ch])                */
    const void* w1;                   /* bf16 [e_ch][c_in + c_in2]  (mode 2)                               */
    const float* bias1;               /* f32 [e_ch]                 (mode 2)                               */
    const float* dw_weight;           /* f32 [ksize*ksize][e_ch]    (tap-major)                            */
    const float* dw_bias;             /* f32 [e_ch]                                                        */
    const void* w2;                   /* bf16 [c_out][e_ch]         (modes 1, 2)                           */
    const float* bias2;               /* f32 [c_out]                (modes 1, 2)                           */
} yms_ms_params;

typedef struct yms_ms_plan yms_ms_plan;       /* opaque: encoded tensor maps + launch geometry */
int yms_ms_plan_create(const yms_ms_params* p, yms_ms_plan** plan);
int yms_ms_plan_run(const yms_ms_plan* plan, void* stream);
int yms_ms_plan_destroy(yms_ms_plan* plan);
/* 2*MACs and algorithmic bytes (inputs + outputs + weights that must cross HBM in this mode). */
int yms_ms_plan_cost(const yms_ms_plan* plan, double* flops, double* bytes);

/* SPPF pooling (components.py:141-146): x1 = maxpool5(x), x2 = maxpool5(x1), x3 = maxpool5(x2)
 * (5x5, stride 1, pad 2), written as channel slots 1..3 of the concat buffer whose slot 0
 * already holds x.  buf: bf16 [B,H,W,4*c] (pixel stride given). */
int yms_sppf_pool(void* buf, int64_t pixel_stride, int batch, int h, int w, int c, void* stream);

/* Nearest x2 upsample (components.py:159-160) written into a channel slice of the concat
 * buffer the next C2f reads (yolov8_neck.py:77-83).  x: bf16 [B,h,w,c] -> y[B,2h,2w,c]. */
int yms_upsample2x(const void* x, int64_t x_pixel_stride, int batch, int h, int w, int c,
                   void* y, int64_t y_pixel_stride, void* stream);

/* ------------------------------------------------------------------------------------------
 * Head decode: replaces the eval branch of Head.forward (yolov8/model/yolov8_head.py:127-144),
 * Head.make_anchors (:146-158) and DFL.forward (components.py:176-191, 16 bins).
 * raw_i: [B, H_i*W_i, 64+nc] channel-last (box logits first), f32 or bf16.
 * pred : f32 [B, A, 4+nc] = (cx, cy, w, h, sigmoid(cls)) in input pixels, A = sum H_i*W_i.
 * If cand_boxes != NULL also emits what the reference post-process derives per anchor
 * (tools/test.py:166-179): xyxy box, best class score, best class (lowest index on ties).
 * ------------------------------------------------------------------------------------------ */
int yms_head_decode(const void* raw0, const void* raw1, const void* raw2, int raw_dtype,
                    int batch, const int32_t* host_hw /* h0,w0,h1,w1,h2,w2 */, int num_classes,
                    const float* host_strides /* 3 */, float* pred,
                    float* cand_boxes /* [B,A,4] or NULL */, float* cand_scores /* [B,A] */,
                    int32_t* cand_labels /* [B,A] */, void* stream);

/* Candidate selection of the reference post-process on an arbitrary prediction tensor
 * (yolov8/tools/test.py:166-179 == tools/train.py:64-72): pred f32 [B,A,4+nc] ->
 * xyxy boxes, per-anchor max class score and argmax (first max). */
int yms_select_candidates(const float* pred, int batch, int anchors, int num_classes,
                          float* boxes, float* scores, int32_t* labels, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched class-aware NMS: replaces the confidence filter + per-class loop around
 * torchvision.ops.nms (yolov8/tools/test.py:181-218 == tools/train.py:74-106).
 * boxes f32 [B,N,4] xyxy, scores f32 [B,N], labels i32 [B,N] in [0, num_classes);
 * n_valid i32 [B] (entries >= n_valid[b] ignored) or NULL.  A box takes part iff
 * score > conf_thr (strict, fp32).  Suppression iff (double)IoU_fp32 > iou_thr between boxes
 * of the same label; output order = label ascending, then score descending, ties by index.
 * keep i32 [B,N]: indices into the N axis (first keep_count[b] valid, rest -1).
 * Limits: N <= 2^20, num_classes <= 2047.
 * ------------------------------------------------------------------------------------------ */
size_t yms_nms_workspace_bytes(int batch, int n);
int yms_nms_batched(const float* boxes, const float* scores, const int32_t* labels,
                    const int32_t* n_valid, int batch, int n, int num_classes,
                    float conf_thr, double iou_thr, int32_t* keep, int32_t* keep_count,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Gather kept detections into padded [B, max_det, 6] = (x1,y1,x2,y2,score,label) rows. */
int yms_gather_detections(const float* boxes, const float* scores, const int32_t* labels,
                          const int32_t* keep, const int32_t* keep_count, int batch, int n,
                          int max_det, float* dets, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YMS_B200_H_ */
