"""Torch-facing operators over the C ABI of libyms_b200.so.

PyTorch is plumbing here (device memory + streams).  Every function launches hand-written
sm_100a kernels on ``torch.cuda.current_stream()``; there is no CPU or eager fallback --
CPU tensors raise ``YmsError``.

The public post-process operators are also registered as torch custom ops under the
``yms_b200::`` namespace (``torch.ops.yms_b200.nms_batched`` etc.).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvParams, DecodeFusion, MsParams, YmsError, check

__all__ = ["ConvPlan", "MsLayerPlan", "stem_conv", "stem_conv_u8", "dwconv", "sppf_pool", "upsample2x", "head_decode",
           "select_candidates", "nms_batched", "gather_detections", "PostBuffers", "YmsError"]


import contextlib

_NULL = contextlib.nullcontext()


def _stream(t: Optional[torch.Tensor] = None) -> int:
    """The current stream of the DEVICE the tensors live on (not of whatever device happens to be current)."""
    return torch.cuda.current_stream(None if t is None else t.device).cuda_stream


def _on(t: torch.Tensor):
    """Context that makes t's device current for a launch (kernels launch on the current device; a stream of another
    device is an error).  No-op in the usual one-process-per-GPU setting."""
    return _NULL if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise YmsError("yolo_ms_b200 operators run on CUDA tensors only (no CPU fallback)")


def _pixel_stride(t: torch.Tensor) -> int:
    """t is a channel-last [..., C] view whose leading dims are contiguous pixels."""
    if t.stride(-1) != 1:
        raise YmsError("channel dimension must be contiguous")
    return t.stride(-2)


class ConvPlan:
    """One convolution unit bound to fixed NHWC bf16 buffers (tensor maps encoded once).

    x, y, x2, residual are [B, H, W, C] views (channel slices of wider buffers allowed).
    weight: bf16 [k*k, c_out, c_in (+ c_in2)], bias: f32 [c_out].
    """

    def __init__(self, x, weight, bias, y, ksize=1, stride=1, act=True, residual=None, x2=None, variant=0):
        _need_cuda(x, weight, bias, y, residual, x2)
        if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16 or bias.dtype != torch.float32:
            raise YmsError("conv: x/weight must be bf16 and bias f32")
        if y.dtype not in (torch.bfloat16, torch.float32):
            raise YmsError("conv: y must be bf16 or f32")
        b, h, w, c_in = x.shape
        c_out = y.shape[-1]
        c_in2 = 0 if x2 is None else x2.shape[-1]
        want = (6, c_out, 64) if variant == 4 else (ksize * ksize, c_out, c_in + c_in2)     # variant 4: pair-packed (include/yms_b200.h)
        if tuple(weight.shape) != want or not weight.is_contiguous():
            raise YmsError(f"conv: weight must be contiguous [{ksize * ksize},{c_out},{c_in + c_in2}], got {tuple(weight.shape)}")
        if tuple(y.shape[:3]) != (b, h // stride, w // stride):
            raise YmsError("conv: output spatial shape mismatch")
        p = ConvParams()
        p.batch, p.in_h, p.in_w, p.c_in, p.c_out = b, h, w, c_in, c_out
        p.ksize, p.stride, p.act = ksize, stride, int(bool(act))
        p.out_dtype = _lib.DTYPE_BF16 if y.dtype == torch.bfloat16 else _lib.DTYPE_F32
        p.c_in2 = c_in2
        p.variant = int(variant)
        self.variant = int(variant)
        p.x, p.x_pixel_stride = x.data_ptr(), _pixel_stride(x)
        if x2 is not None:
            p.x2, p.x2_pixel_stride = x2.data_ptr(), _pixel_stride(x2)
        p.y, p.y_pixel_stride = y.data_ptr(), _pixel_stride(y)
        if residual is not None:
            p.residual, p.res_pixel_stride = residual.data_ptr(), _pixel_stride(residual)
        p.weight, p.bias = weight.data_ptr(), bias.data_ptr()
        self._keep = (x, weight, bias, y, residual, x2)      # keep the storages alive
        self._h = C.c_void_p()
        self._lib = _lib.load()
        with _on(y):
            check(self._lib.yms_conv_plan_create(C.byref(p), C.byref(self._h)), "yms_conv_plan_create")
        fl, by = C.c_double(), C.c_double()
        self._lib.yms_conv_plan_cost(self._h, C.byref(fl), C.byref(by))
        self.flops, self.bytes = fl.value, by.value
        self.desc = (f"conv{ksize}x{ksize}/s{stride} {c_in}{'+' + str(c_in2) if c_in2 else ''}->{c_out} @{h}x{w}"
                     f"{' +res' if residual is not None else ''}{' f32' if y.dtype == torch.float32 else ''}")

    def fuse_decode(self, branch: str, stride: torch.Tensor, pred: torch.Tensor, anchor_base: int,
                    cand_boxes: Optional[torch.Tensor] = None, cand_scores: Optional[torch.Tensor] = None,
                    cand_labels: Optional[torch.Tensor] = None) -> None:
        """Turn the plan of a head branch's final biased 1x1 conv (f32 output, act=False) into one that decodes in its
        epilogue instead of storing logits (yms_conv_plan_fuse_decode).  branch: 'box' | 'cls'; stride: DEVICE f32 [1]
        view read at run time; pred f32 [B, A, 4+nc]; candidates f32 [B,A,4] / f32 [B,A] / i32 [B,A]."""
        _need_cuda(stride, pred, cand_boxes, cand_scores, cand_labels)
        if pred.dtype != torch.float32 or not pred.is_contiguous() or pred.dim() != 3 or stride.dtype != torch.float32:
            raise YmsError("fuse_decode: pred must be contiguous f32 [B,A,4+nc] and stride f32")
        for t, dt in ((cand_boxes, torch.float32), (cand_scores, torch.float32), (cand_labels, torch.int32)):
            if t is not None and (t.dtype != dt or not t.is_contiguous()):
                raise YmsError("fuse_decode: candidate buffers must be contiguous f32 / f32 / int32")
        x = self._keep[0]
        f = DecodeFusion()
        f.branch = {"box": 1, "cls": 2}[branch]
        f.map_h, f.map_w = x.shape[1], x.shape[2]
        f.anchor_base, f.anchors, f.num_classes = int(anchor_base), pred.shape[1], pred.shape[2] - 4
        f.stride, f.pred = stride.data_ptr(), pred.data_ptr()
        f.cand_boxes = None if cand_boxes is None else cand_boxes.data_ptr()
        f.cand_scores = None if cand_scores is None else cand_scores.data_ptr()
        f.cand_labels = None if cand_labels is None else cand_labels.data_ptr()
        if pred.shape[0] != x.shape[0]:
            raise YmsError("fuse_decode: batch mismatch")
        check(self._lib.yms_conv_plan_fuse_decode(self._h, C.byref(f)), "yms_conv_plan_fuse_decode")
        self._keep = self._keep + (stride, pred, cand_boxes, cand_scores, cand_labels)
        fl, by = C.c_double(), C.c_double()
        self._lib.yms_conv_plan_cost(self._h, C.byref(fl), C.byref(by))
        self.flops, self.bytes = fl.value, by.value
        self.desc = self.desc.replace(" f32", "") + f" +decode({branch})"

    def add_upsampled(self, t: torch.Tensor) -> None:
        """Add nearest-upsampled fp32 partial sums t [B, H/2, W/2, c_out] to the accumulator before bias / activation
        (yms_conv_plan_add_upsampled: the upsample + concat half of a 1x1 conv over cat[upsample2x(a), b])."""
        _need_cuda(t)
        x, y = self._keep[0], self._keep[3]
        b, h, w, _ = x.shape
        if t.dtype != torch.float32 or tuple(t.shape) != (b, h // 2, w // 2, y.shape[-1]):
            raise YmsError(f"add_upsampled: t must be f32 [{b},{h // 2},{w // 2},{y.shape[-1]}], got {tuple(t.shape)} {t.dtype}")
        check(self._lib.yms_conv_plan_add_upsampled(self._h, t.data_ptr(), _pixel_stride(t), h, w), "yms_conv_plan_add_upsampled")
        self._keep = self._keep + (t,)
        fl, by = C.c_double(), C.c_double()
        self._lib.yms_conv_plan_cost(self._h, C.byref(fl), C.byref(by))
        self.flops, self.bytes = fl.value, by.value
        self.desc += " +up(f32)"

    def run(self) -> None:
        y = self._keep[3]
        with _on(y):
            check(self._lib.yms_conv_plan_run(self._h, _stream(y)), "yms_conv_plan_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.yms_conv_plan_destroy(h)
            except Exception:
                pass
            self._h = None


class MsLayerPlan:
    """One MS-Block branch layer  pw1 (1x1) -> depthwise k x k -> pw2 (1x1)  bound to fixed NHWC bf16 buffers
    (yms_ms_plan_create).  mode 0: depthwise only (e -> y); mode 1: depthwise -> pw2 (e -> y); mode 2: the whole layer
    (x [+ x2] -> y), the expanded tensors never reach HBM.

    dw_weight f32 [k*k, E], dw_bias f32 [E]; w1 bf16 [E, c_in (+ c_in2)], bias1 f32 [E]; w2 bf16 [c_out, E], bias2 f32 [c_out].
    Raises YmsError (YMS_E_UNSUPPORTED) when the mode does not fit the hardware budgets -- callers fall back to a lower mode."""

    def __init__(self, mode, y, ksize, dw_weight, dw_bias, e=None, x=None, x2=None, w1=None, bias1=None, w2=None, bias2=None,
                 act2=True):
        _need_cuda(y, dw_weight, dw_bias, e, x, x2, w1, bias1, w2, bias2)
        src = e if mode != 2 else x
        if src is None or src.dtype != torch.bfloat16 or y.dtype != torch.bfloat16:
            raise YmsError("ms layer: activations must be bf16")
        b, h, w, _ = src.shape
        e_ch = dw_weight.shape[1]
        if tuple(dw_weight.shape) != (ksize * ksize, e_ch) or dw_weight.dtype != torch.float32 or not dw_weight.is_contiguous() \
                or dw_bias.dtype != torch.float32 or tuple(dw_bias.shape) != (e_ch,):
            raise YmsError("ms layer: dw_weight must be contiguous f32 [k*k, E] and dw_bias f32 [E]")
        if tuple(y.shape[:3]) != (b, h, w):
            raise YmsError("ms layer: output spatial shape mismatch")
        p = MsParams()
        p.mode, p.batch, p.h, p.w, p.ksize, p.e_ch = int(mode), b, h, w, int(ksize), e_ch
        p.act2 = int(bool(act2))
        p.y, p.y_pixel_stride = y.data_ptr(), _pixel_stride(y)
        p.dw_weight, p.dw_bias = dw_weight.data_ptr(), dw_bias.data_ptr()
        if mode != 2:
            if e.shape[-1] != e_ch:
                raise YmsError("ms layer: e must have E channels")
            p.e, p.e_pixel_stride = e.data_ptr(), _pixel_stride(e)
        else:
            c_in, c_in2 = x.shape[-1], (0 if x2 is None else x2.shape[-1])
            if w1 is None or w1.dtype != torch.bfloat16 or tuple(w1.shape) != (e_ch, c_in + c_in2) or not w1.is_contiguous() \
                    or bias1 is None or bias1.dtype != torch.float32:
                raise YmsError(f"ms layer: w1 must be contiguous bf16 [{e_ch},{c_in + c_in2}] and bias1 f32")
            p.c_in, p.c_in2 = c_in, c_in2
            p.x, p.x_pixel_stride = x.data_ptr(), _pixel_stride(x)
            if x2 is not None:
                p.x2, p.x2_pixel_stride = x2.data_ptr(), _pixel_stride(x2)
            p.w1, p.bias1 = w1.data_ptr(), bias1.data_ptr()
        if mode >= 1:
            c_out = y.shape[-1]
            if w2 is None or w2.dtype != torch.bfloat16 or tuple(w2.shape) != (c_out, e_ch) or not w2.is_contiguous() \
                    or bias2 is None or bias2.dtype != torch.float32:
                raise YmsError(f"ms layer: w2 must be contiguous bf16 [{c_out},{e_ch}] and bias2 f32")
            p.c_out = c_out
            p.w2, p.bias2 = w2.data_ptr(), bias2.data_ptr()
        elif y.shape[-1] != e_ch:
            raise YmsError("ms layer: mode 0 output must have E channels")
        self._keep = (y, dw_weight, dw_bias, e, x, x2, w1, bias1, w2, bias2)
        self.mode = int(mode)
        self._h = C.c_void_p()
        self._lib = _lib.load()
        with _on(y):
            check(self._lib.yms_ms_plan_create(C.byref(p), C.byref(self._h)), "yms_ms_plan_create")
        fl, by = C.c_double(), C.c_double()
        self._lib.yms_ms_plan_cost(self._h, C.byref(fl), C.byref(by))
        self.flops, self.bytes = fl.value, by.value
        what = {0: f"dw{ksize}x{ksize} {e_ch}", 1: f"dw{ksize}x{ksize} {e_ch} ->pw2 {y.shape[-1]}",
                2: f"ms-layer {'' if x is None else x.shape[-1]}{'' if x2 is None else '+' + str(x2.shape[-1])}->{e_ch} dw{ksize}x{ksize} ->{y.shape[-1]}"}
        self.desc = f"{what[self.mode]} @{h}x{w}"

    def run(self) -> None:
        y = self._keep[0]
        with _on(y):
            check(self._lib.yms_ms_plan_run(self._h, _stream(y)), "yms_ms_plan_run")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.yms_ms_plan_destroy(h)
            except Exception:
                pass
            self._h = None


def stem_conv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, y: torch.Tensor) -> None:
    """x f32 NCHW [B,3,H,W]; weight f32 [c_out,3,3,3] (BN folded); y bf16 [B,H/2,W/2,c_out]."""
    _need_cuda(x, weight, bias, y)
    if x.dtype != torch.float32 or not x.is_contiguous() or x.shape[1] != 3:
        raise YmsError("stem: x must be contiguous f32 [B,3,H,W]")
    b, _, h, w = x.shape
    with _on(y):
        check(_lib.load().yms_stem_conv(x.data_ptr(), b, h, w, y.shape[-1], weight.data_ptr(), bias.data_ptr(),
                                        y.data_ptr(), _pixel_stride(y), _stream(y)), "yms_stem_conv")


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def stem_conv_u8(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, y: torch.Tensor,
                 mean=IMAGENET_MEAN, std=IMAGENET_STD) -> None:
    """x uint8 NHWC [B,H,W,3]; ToTensor + Normalize(mean, std) fused into the stem (tools/test.py:114-119)."""
    _need_cuda(x, weight, bias, y)
    if x.dtype != torch.uint8 or not x.is_contiguous() or x.dim() != 4 or x.shape[3] != 3:
        raise YmsError("stem_u8: x must be contiguous uint8 [B,H,W,3]")
    b, h, w, _ = x.shape
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with _on(y):
        check(_lib.load().yms_stem_conv_u8(x.data_ptr(), b, h, w, y.shape[-1], weight.data_ptr(), bias.data_ptr(), m, s,
                                           y.data_ptr(), _pixel_stride(y), _stream(y)), "yms_stem_conv_u8")


def dwconv(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, y: torch.Tensor, ksize: int) -> None:
    """Depthwise kxk + bias + SiLU.  x,y bf16 [B,H,W,C]; weight f32 [k*k, C]; bias f32 [C]."""
    _need_cuda(x, weight, bias, y)
    b, h, w, c = x.shape
    with _on(y):
        check(_lib.load().yms_dwconv(x.data_ptr(), _pixel_stride(x), b, h, w, c, ksize, weight.data_ptr(), bias.data_ptr(),
                                     y.data_ptr(), _pixel_stride(y), _stream(y)), "yms_dwconv")


def sppf_pool(buf: torch.Tensor, c: int) -> None:
    """buf bf16 [B,H,W,>=4c]: slot 0 holds x; writes the 5/9/13 max pools into slots 1..3."""
    _need_cuda(buf)
    b, h, w, _ = buf.shape
    with _on(buf):
        check(_lib.load().yms_sppf_pool(buf.data_ptr(), _pixel_stride(buf), b, h, w, c, _stream(buf)), "yms_sppf_pool")


def upsample2x(x: torch.Tensor, y: torch.Tensor) -> None:
    _need_cuda(x, y)
    b, h, w, c = x.shape
    with _on(y):
        check(_lib.load().yms_upsample2x(x.data_ptr(), _pixel_stride(x), b, h, w, c, y.data_ptr(), _pixel_stride(y),
                                         _stream(y)), "yms_upsample2x")


def head_decode(raw: Sequence[torch.Tensor], strides: Sequence[float], num_classes: int,
                with_candidates: bool = False):
    """raw: 3 tensors [B,H_i,W_i,64+nc] (channel-last, contiguous, f32 or bf16).
    Returns pred [B,A,4+nc] f32 (and (boxes_xyxy, scores, labels) when with_candidates)."""
    _need_cuda(*raw)
    if len(raw) != 3:
        raise YmsError("head_decode expects 3 scales")
    dt = raw[0].dtype
    if dt not in (torch.float32, torch.bfloat16) or any(r.dtype != dt or not r.is_contiguous() for r in raw):
        raise YmsError("head_decode: raw tensors must be contiguous and share dtype f32/bf16")
    b = raw[0].shape[0]
    no = 64 + num_classes
    hw = (C.c_int32 * 6)()
    a = 0
    for i, r in enumerate(raw):
        if r.shape[0] != b or r.shape[-1] != no:
            raise YmsError("head_decode: bad raw shape")
        hw[2 * i], hw[2 * i + 1] = r.shape[1], r.shape[2]
        a += r.shape[1] * r.shape[2]
    st = (C.c_float * 3)(*[float(s) for s in strides])
    dev = raw[0].device
    pred = torch.empty((b, a, 4 + num_classes), dtype=torch.float32, device=dev)
    if with_candidates:
        boxes = torch.empty((b, a, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((b, a), dtype=torch.float32, device=dev)
        labels = torch.empty((b, a), dtype=torch.int32, device=dev)
        ptrs = (boxes.data_ptr(), scores.data_ptr(), labels.data_ptr())
    else:
        boxes = scores = labels = None
        ptrs = (None, None, None)
    with _on(pred):
        check(_lib.load().yms_head_decode(raw[0].data_ptr(), raw[1].data_ptr(), raw[2].data_ptr(),
                                          _lib.DTYPE_F32 if dt == torch.float32 else _lib.DTYPE_BF16, b, hw, num_classes, st,
                                          pred.data_ptr(), *ptrs, _stream(pred)), "yms_head_decode")
    return (pred, (boxes, scores, labels)) if with_candidates else pred


def select_candidates(pred: torch.Tensor):
    """pred f32 [B,A,4+nc] -> (xyxy [B,A,4], best score [B,A], best class [B,A] int32)."""
    _need_cuda(pred)
    if pred.dtype != torch.float32 or pred.dim() != 3:
        raise YmsError("select_candidates: pred must be f32 [B,A,4+nc]")
    pred = pred.contiguous()
    b, a, no = pred.shape
    boxes = torch.empty((b, a, 4), dtype=torch.float32, device=pred.device)
    scores = torch.empty((b, a), dtype=torch.float32, device=pred.device)
    labels = torch.empty((b, a), dtype=torch.int32, device=pred.device)
    with _on(pred):
        check(_lib.load().yms_select_candidates(pred.data_ptr(), b, a, no - 4, boxes.data_ptr(), scores.data_ptr(),
                                                labels.data_ptr(), _stream(pred)), "yms_select_candidates")
    return boxes, scores, labels


class PostBuffers:
    """Pre-allocated outputs + workspace of the post-process for a fixed [B, N] candidate shape: keep / count (yms_nms_batched),
    its workspace, and optionally the padded detection rows (yms_gather_detections).  Lets a whole step be captured in ONE
    CUDA graph with zero allocations per call."""

    def __init__(self, batch: int, n: int, device, max_det: Optional[int] = None):
        lib = _lib.load()
        self.keep = torch.empty((batch, n), dtype=torch.int32, device=device)
        self.count = torch.empty((batch,), dtype=torch.int32, device=device)
        self.ws_bytes = lib.yms_nms_workspace_bytes(batch, n)
        self.ws = torch.empty((max(self.ws_bytes, 8),), dtype=torch.uint8, device=device)
        self.dets = None if max_det is None else torch.empty((batch, max_det, 6), dtype=torch.float32, device=device)


def _f32(v: float) -> float:
    """float(np.float32(v)): the reference compares fp32 scores with the Python scalar cast to fp32 (no device work)."""
    return C.c_float(v).value


def nms_batched(boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, conf_thr: float, iou_thr: float,
                num_classes: int, n_valid: Optional[torch.Tensor] = None, out: Optional[PostBuffers] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """boxes f32 [B,N,4] xyxy, scores f32 [B,N], labels int32 [B,N].
    Returns (keep int32 [B,N] padded with -1, keep_count int32 [B]); with `out` nothing is allocated."""
    _need_cuda(boxes, scores, labels, n_valid)
    if boxes.dtype != torch.float32 or scores.dtype != torch.float32 or labels.dtype != torch.int32:
        raise YmsError("nms_batched: dtypes must be f32/f32/int32")
    boxes, scores, labels = boxes.contiguous(), scores.contiguous(), labels.contiguous()
    b, n = scores.shape
    if out is None:
        out = PostBuffers(b, n, boxes.device)
    elif tuple(out.keep.shape) != (b, n):
        raise YmsError("nms_batched: `out` was allocated for another shape")
    nv = None
    if n_valid is not None:
        nv = n_valid.to(torch.int32).contiguous()
    with _on(boxes):
        check(_lib.load().yms_nms_batched(boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(),
                                          None if nv is None else nv.data_ptr(), b, n, num_classes, _f32(conf_thr), float(iou_thr),
                                          out.keep.data_ptr(), out.count.data_ptr(), out.ws.data_ptr(), out.ws_bytes, _stream(boxes)),
              "yms_nms_batched")
    return out.keep, out.count


def gather_detections(boxes, scores, labels, keep, count, max_det: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """-> f32 [B, max_det, 6] rows (x1,y1,x2,y2,score,label); unused rows have label -1."""
    _need_cuda(boxes, scores, labels, keep, count)
    b, n = scores.shape
    dets = out if out is not None else torch.empty((b, max_det, 6), dtype=torch.float32, device=boxes.device)
    if tuple(dets.shape) != (b, max_det, 6) or dets.dtype != torch.float32 or not dets.is_contiguous():
        raise YmsError("gather_detections: `out` must be contiguous f32 [B, max_det, 6]")
    with _on(dets):
        check(_lib.load().yms_gather_detections(boxes.data_ptr(), scores.data_ptr(), labels.data_ptr(), keep.data_ptr(),
                                                count.data_ptr(), b, n, max_det, dets.data_ptr(), _stream(dets)),
              "yms_gather_detections")
    return dets


# ---------------------------------------------------------------------------------------------
# torch custom-op registration (thin shims over the functions above)
# ---------------------------------------------------------------------------------------------
def _register():
    try:
        from torch.library import custom_op
    except Exception:  # pragma: no cover
        return

    @custom_op("yms_b200::nms_batched", mutates_args=(), device_types="cuda")
    def _nms(boxes: torch.Tensor, scores: torch.Tensor, labels: torch.Tensor, conf_thr: float, iou_thr: float,
             num_classes: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return nms_batched(boxes, scores, labels, conf_thr, iou_thr, num_classes)

    @custom_op("yms_b200::select_candidates", mutates_args=(), device_types="cuda")
    def _sel(pred: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return select_candidates(pred)

    @custom_op("yms_b200::head_decode", mutates_args=(), device_types="cuda")
    def _dec(raw0: torch.Tensor, raw1: torch.Tensor, raw2: torch.Tensor, strides: Sequence[float],
             num_classes: int) -> torch.Tensor:
        return head_decode((raw0, raw1, raw2), strides, num_classes)


_register()
