"""Drop-in module classes: same constructor / forward API and the same state_dict keys as the
reference's ``yolov8`` package, executed by hand-written sm_100a kernels.

Reference API mirrored here (paths relative to rafaelghiorzi/YOLO-MS):
  YOLOv8(version, num_classes, dfl_ch=16)            yolov8/yolov8.py:8-31
  Backbone(version, in_channels=3, shortcut=True)    yolov8/model/yolov8_backbone.py:34-74
  Neck(version)                                      yolov8/model/yolov8_neck.py:55-94
  Head(version, ch=16, num_classes=80), .stride      yolov8/model/yolov8_head.py:73-158
  Conv / Bottleneck / C2f / SPPF / Upsample / DFL / yolo_params
                                                     yolov8/model/components.py:69-209
plus the repo-local ``MSBlock`` (``block='ms'``), which the reference only sketches
(annotations.md:66-133).

The modules own ordinary ``nn.Parameter``s / BN buffers (so ``state_dict`` round-trips with
reference checkpoints); ``forward`` compiles, per input shape, a static launch program
(engine.Program) from the *current* parameter values with BN folded, and replays it.  There is
no CPU path and no autograd: a CPU input raises ``YmsError``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import ops
from ._lib import YmsError
from .engine import Program, fold_conv_bn, nchw_f32_to_nhwc_bf16, nhwc_to_nchw_f32, pack_weight

FUSE_DECODE = os.environ.get("YMS_FUSE_DECODE", "1") != "0"    # decode in the epilogue of the head's final convs (YOLOv8 programs)

MS_FUSE = int(os.environ.get("YMS_MS_FUSE", "2"))               # MS-Block layers: 2 = pw1 -> depthwise -> pw2 in one kernel where it
                                                                # fits, 1 = depthwise -> pw2, 0 = three launches

OVERLAP_HEAD = os.environ.get("YMS_OVERLAP_HEAD", "1") != "0"    # head scales 0 / 1 on graph branches next to the rest of the neck

FUSE_UPSAMPLE = os.environ.get("YMS_FUSE_UPSAMPLE", "1") != "0"  # neck: upsample + concat inside the consumer C2f's first 1x1 conv

_VERSIONS = {  # depth, width, ratio  (components.py:193-209)
    "n": (1 / 3, 1 / 4, 2.0), "s": (1 / 3, 1 / 2, 2.0), "m": (2 / 3, 3 / 4, 1.5),
    "l": (1.0, 1.0, 1.0), "x": (1.0, 1.25, 1.0),
}


def yolo_params(version):
    if version not in _VERSIONS:
        raise ValueError(f"Unknown YOLOv8 version: {version}")
    return _VERSIONS[version]


_WEIGHTS_EPOCH = [0]     # bumped whenever ANY compiled module's parameters may have been replaced (load_state_dict / .to() / ...)


class _Compiled(nn.Module):
    """Mixin: per-shape program cache.  Programs bake in folded-BN bf16 weights, so they are dropped whenever parameters may
    have changed: `load_state_dict` / `_apply` (.to, .half, ...) on this module OR on any other compiled module -- a parent's
    program holds its children's weights, and `model.head.load_state_dict(...)` must invalidate `model`'s program too, hence
    one process-wide epoch instead of per-module flags.  In-place edits that bypass those calls (`p.data.copy_`, an optimizer
    step) cannot be seen from here: call `refresh()` after them."""

    def _programs(self) -> Dict:
        cache = self.__dict__.get("_yms_cache")
        if cache is None or self.__dict__.get("_yms_epoch") != _WEIGHTS_EPOCH[0]:
            cache = {}
            self.__dict__["_yms_cache"] = cache
            self.__dict__["_yms_epoch"] = _WEIGHTS_EPOCH[0]
        return cache

    def refresh(self):
        """Drop compiled programs (call after modifying parameters in place)."""
        _WEIGHTS_EPOCH[0] += 1
        for m in self.modules():
            m.__dict__.pop("_yms_cache", None)
        return self

    def _apply(self, fn, *a, **k):
        self.refresh()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.refresh()
        return super().load_state_dict(*a, **k)


def img_device(image) -> torch.device:
    return image.tensor.device


def _device_of(m: nn.Module) -> torch.device:
    return next(m.parameters()).device


# =============================================================================================
# building blocks
# =============================================================================================
class Conv(_Compiled):
    """conv(bias=False) + BN(eps 1e-3, momentum 0.03) + SiLU  (components.py:69-77)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, groups=1, activation=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=False, groups=groups)
        self.bn = nn.BatchNorm2d(out_channels, eps=0.001, momentum=0.03)
        self.activation = nn.SiLU(inplace=True) if activation else nn.Identity()

    # ---- weight preparation --------------------------------------------------------------
    def folded(self):
        return fold_conv_bn(self.conv.weight.detach(), self.bn.weight.detach(), self.bn.bias.detach(),
                            self.bn.running_mean, self.bn.running_var, self.bn.eps)

    def dw_folded(self):
        """Depthwise unit: (f32 [k*k, C] tap-major weights, f32 [C] bias) with BN folded."""
        w, b = self.folded()
        co, k = self.conv.out_channels, self.conv.kernel_size[0]
        return w.reshape(co, k * k).t().contiguous(), b.contiguous()

    @property
    def has_act(self):
        return isinstance(self.activation, nn.SiLU)

    def emit(self, P: Program, x, out=None, residual=None, x2=None, dup_k=False):
        """x: NHWC bf16 view.  dup_k: weights repeated over both sources (conv(x + x2))."""
        conv = self.conv
        k, s = conv.kernel_size[0], conv.stride[0]
        if conv.padding[0] != k // 2:
            raise YmsError("Conv: only padding = kernel_size // 2 is implemented")
        w, b = self.folded()
        bsz, h, wd, _ = x.shape
        co = conv.out_channels
        if out is None:
            out = P.buf(bsz, h // s, wd // s, co)
        if conv.groups == 1:
            if dup_k:
                w = torch.cat([w, w], 1)
            P.conv(pack_weight(w), b.contiguous(), x, out, ksize=k, stride=s, act=self.has_act, residual=residual, x2=x2)
        elif conv.groups == conv.in_channels == co and s == 1 and self.has_act and residual is None and x2 is None:
            wk, bb = self.dw_folded()
            P.hold(wk, bb)
            P.ms_layer(ops.MsLayerPlan(0, out, k, wk, bb, e=x))
        else:
            raise YmsError("Conv: only groups=1 or depthwise stride-1 convolutions are implemented")
        return out

    def forward(self, x):
        return _run_standalone(self, (x,), lambda P, xs: (self.emit(P, xs[0]),))[0]


class Bottleneck(_Compiled):
    """Two 3x3 Conv units + residual (components.py:80-93)."""

    def __init__(self, in_channels, out_channels, shortcut=True):
        super().__init__()
        self.conv1 = Conv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.conv2 = Conv(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.shortcut = shortcut

    def emit(self, P, x, out=None):
        t = self.conv1.emit(P, x)
        return self.conv2.emit(P, t, out=out, residual=x if self.shortcut else None)

    def forward(self, x):
        return _run_standalone(self, (x,), lambda P, xs: (self.emit(P, xs[0]),))[0]


class C2f(_Compiled):
    """components.py:96-122.  Concat order [b_n, ..., b_1, x1, x2]; the FIRST half is chained;
    bottlenecks always keep their shortcut (the reference never forwards `shortcut`)."""

    def __init__(self, in_channels, out_channels, num_bottlenecks, shortcut=True):
        super().__init__()
        self.mid_channels = out_channels // 2
        self.num_bottlenecks = num_bottlenecks
        self.conv1 = Conv(in_channels, out_channels, kernel_size=1, stride=1, padding=0)
        self.m = nn.ModuleList([Bottleneck(self.mid_channels, self.mid_channels) for _ in range(num_bottlenecks)])
        self.conv2 = Conv((num_bottlenecks + 2) * out_channels // 2, out_channels, kernel_size=1, stride=1, padding=0)

    def emit(self, P, x, out=None, up_src=None):
        """up_src (neck only): x is the concat buffer [upsample2x(up_src) | skip] whose FIRST slot has NOT been written; conv1
        (1x1, linear per pixel) is then split as  upsample2x(W_a . up_src) + W_b . skip : the first term runs at half resolution
        as a linear fp32 1x1 conv, the second adds it in its epilogue (yms_conv_plan_add_upsampled) -- no upsample kernel, no
        upsampled tensor, 4x fewer MACs for that half (components.py:159-160 + yolov8_neck.py:77-83 + components.py:108)."""
        n, h = self.num_bottlenecks, self.mid_channels
        b, hh, ww, _ = x.shape
        cat = P.buf(b, hh, ww, (n + 2) * h)
        if up_src is not None:
            c_low = up_src.shape[-1]
            wf, bf = self.conv1.folded()
            co = wf.shape[0]
            part = P.buf(b, hh // 2, ww // 2, co, dtype=torch.float32)
            P.conv(pack_weight(wf[:, :c_low].contiguous()), torch.zeros_like(bf), up_src, part, ksize=1, stride=1, act=False)
            P.conv(pack_weight(wf[:, c_low:].contiguous()), bf.contiguous(), x[..., c_low:], cat[..., n * h:(n + 2) * h], ksize=1, stride=1,
                   act=self.conv1.has_act, up_add=part)
        else:
            self.conv1.emit(P, x, out=cat[..., n * h:(n + 2) * h])
        cur = cat[..., n * h:(n + 1) * h]
        for j, blk in enumerate(self.m):
            dst = cat[..., (n - 1 - j) * h:(n - j) * h]
            blk.emit(P, cur, out=dst)
            cur = dst
        return self.conv2.emit(P, cat, out=out)

    def forward(self, x):
        return _run_standalone(self, (x,), lambda P, xs: (self.emit(P, xs[0]),))[0]


class MSBlock(_Compiled):
    """Repo-local MS-Block (NOT in the reference; after arXiv 2308.05480).  in_conv 1x1
    (C_in -> 3c, c = C_out/2); branch 0 identity; branch i>=1: (x_i + y_{i-1}) -> L x
    [1x1 c->2c, depthwise kxk, 1x1 2c->c]; concat; out_conv 1x1 (3c -> C_out)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, layers_num=1):
        super().__init__()
        c = out_channels // 2
        self.mid_channels = c
        self.kernel_size = kernel_size
        self.in_conv = Conv(in_channels, 3 * c, kernel_size=1, stride=1, padding=0)
        branches = []
        for _ in range(2):
            layers = []
            for _ in range(layers_num):
                layer = nn.Module()
                layer.pw1 = Conv(c, 2 * c, kernel_size=1, stride=1, padding=0)
                layer.dw = Conv(2 * c, 2 * c, kernel_size=kernel_size, stride=1, padding=kernel_size // 2, groups=2 * c)
                layer.pw2 = Conv(2 * c, c, kernel_size=1, stride=1, padding=0)
                layers.append(layer)
            branches.append(nn.ModuleList(layers))
        self.branches = nn.ModuleList(branches)
        self.out_conv = Conv(3 * c, out_channels, kernel_size=1, stride=1, padding=0)

    def _emit_layer(self, P, layer, src, dup, dst):
        """One branch layer pw1 -> dw -> pw2 over src.  dup: src is the 2c-wide slice [a | b] of the two tensors whose SUM the
        layer takes -- conv(a + b) = K-concatenated GEMM with the pw1 weights repeated along K (exact in fp32 accumulation).
        Three implementations (csrc/ms_fused.cu): the whole layer in one kernel (when it fits the shared-memory / TMEM
        budgets), pw1 + (dw -> pw2), or three launches; the fastest one on this layer's buffers is kept (engine.pick_fastest;
        YMS_AUTOTUNE=0: the most fused one that fits).  They differ in fp32 summation order only."""
        k = self.kernel_size
        b, hh, ww, _ = src.shape
        if not (layer.pw1.has_act and layer.dw.has_act and layer.pw1.conv.groups == 1 and layer.pw2.conv.groups == 1):
            e = layer.pw1.emit(P, src, dup_k=dup)
            return layer.pw2.emit(P, layer.dw.emit(P, e), out=dst)
        wd, bd = layer.dw.dw_folded()
        w1f, b1 = layer.pw1.folded()
        w2f, b2 = layer.pw2.folded()
        if dup:
            w1f = torch.cat([w1f, w1f], 1)
        w1 = w1f.reshape(w1f.shape[0], -1).contiguous().to(torch.bfloat16)        # [E, K]
        w2 = w2f.reshape(w2f.shape[0], -1).contiguous().to(torch.bfloat16)        # [c, E]
        b1, b2 = b1.contiguous(), b2.contiguous()
        e_ch = w1.shape[0]
        if dst is None:
            dst = P.buf(b, hh, ww, w2.shape[0])
        act2 = layer.pw2.has_act
        cands = []
        if MS_FUSE >= 2:
            try:
                cands.append(("fused", [ops.MsLayerPlan(2, dst, k, wd, bd, x=src, w1=w1, bias1=b1, w2=w2, bias2=b2, act2=act2)]))
            except YmsError:
                pass
        e = torch.empty((b, hh, ww, e_ch), dtype=torch.bfloat16, device=src.device)
        pw1 = ops.ConvPlan(src, w1.unsqueeze(0), b1, e, ksize=1, stride=1, act=True)
        if MS_FUSE >= 1:
            try:
                cands.append(("dw+pw2", [pw1, ops.MsLayerPlan(1, dst, k, wd, bd, e=e, w2=w2, bias2=b2, act2=act2)]))
            except YmsError:
                pass
        d = torch.empty_like(e)
        cands.append(("unfused", [pw1, ops.MsLayerPlan(0, d, k, wd, bd, e=e), ops.ConvPlan(d, w2.unsqueeze(0), b2, dst, ksize=1, stride=1, act=act2)]))
        key = ("ms-layer", str(src.device), tuple(src.shape), src.stride(-2), dst.stride(-2), e_ch, k, tuple(c[0] for c in cands))
        tag, plans = P.pick_fastest(key, cands)
        del cands, e, d                                     # the losing candidates' scratch tensors die with their plans
        for pl in plans:
            P.ms_layer(pl)
        return dst

    def emit(self, P, x, out=None, up_src=None):
        """One buffer [x0 | x1 | x2 | y1 | y2] (5c channels): in_conv fills the first three slots, the branches the last two.
        The branch inputs x1 + x0 and x2 + y1 are then ADJACENT slices ([x0|x1], [x2|y1]): one 2c-wide source each, and
        out_conv reads cat[x0, y1, y2] as x0 plus the adjacent pair [y1|y2].
        up_src (neck only): as in C2f.emit -- x is [upsample2x(up_src) | skip] with the first slot NOT written, and in_conv is
        split into a half-resolution fp32 partial-sum conv over up_src plus a conv over the skip that adds it in its epilogue."""
        c = self.mid_channels
        b, hh, ww, _ = x.shape
        y = P.buf(b, hh, ww, 5 * c)
        if up_src is not None:
            c_low = up_src.shape[-1]
            wf, bf = self.in_conv.folded()
            part = P.buf(b, hh // 2, ww // 2, wf.shape[0], dtype=torch.float32)
            P.conv(pack_weight(wf[:, :c_low].contiguous()), torch.zeros_like(bf), up_src, part, ksize=1, stride=1, act=False)
            P.conv(pack_weight(wf[:, c_low:].contiguous()), bf.contiguous(), x[..., c_low:], y[..., :3 * c], ksize=1, stride=1,
                   act=self.in_conv.has_act, up_add=part)
        else:
            self.in_conv.emit(P, x, out=y[..., :3 * c])
        for bi, layers in enumerate(self.branches):
            t = y[..., 0:2 * c] if bi == 0 else y[..., 2 * c:4 * c]
            for li, layer in enumerate(layers):
                dst = y[..., (3 + bi) * c:(4 + bi) * c] if li == len(layers) - 1 else None
                t = self._emit_layer(P, layer, t, li == 0, dst)
        return self.out_conv.emit(P, y[..., :c], out=out, x2=y[..., 3 * c:5 * c])

    def forward(self, x):
        return _run_standalone(self, (x,), lambda P, xs: (self.emit(P, xs[0]),))[0]


class SPPF(_Compiled):
    """components.py:125-150."""

    def __init__(self, in_channels, out_channels, kernel_size=5):
        super().__init__()
        if kernel_size != 5:
            raise YmsError("SPPF: only kernel_size=5 is implemented")
        hidden = in_channels // 2
        self.conv1 = Conv(in_channels, hidden, kernel_size=1, stride=1, padding=0)
        self.conv2 = Conv(hidden * 4, out_channels, kernel_size=1, stride=1, padding=0)
        self.m = nn.MaxPool2d(kernel_size=kernel_size, stride=1, padding=kernel_size // 2)

    def emit(self, P, x, out=None):
        b, hh, ww, _ = x.shape
        hc = self.conv1.conv.out_channels
        cat = P.buf(b, hh, ww, 4 * hc)
        self.conv1.emit(P, x, out=cat[..., :hc])
        P.add(lambda: ops.sppf_pool(cat, hc), nbytes=2.0 * 4 * b * hh * ww * hc, name=f"sppf_pool {hc} @{hh}x{ww}")
        return self.conv2.emit(P, cat, out=out)

    def forward(self, x):
        return _run_standalone(self, (x,), lambda P, xs: (self.emit(P, xs[0]),))[0]


class Upsample(nn.Module):
    """Nearest x2 (components.py:153-160)."""

    def __init__(self, scale_factor=2, mode="nearest"):
        super().__init__()
        if scale_factor != 2 or mode != "nearest":
            raise YmsError("Upsample: only nearest x2 is implemented")
        self.scale_factor, self.mode = scale_factor, mode

    def emit(self, P, x, out):
        P.add(lambda: ops.upsample2x(x, out), nbytes=2.0 * 5 * x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3],
              name=f"upsample2x {x.shape[3]} @{x.shape[1]}x{x.shape[2]}")
        return out

    def forward(self, x):
        xin = nchw_f32_to_nhwc_bf16(x)
        b, h, w, c = xin.shape
        y = torch.empty((b, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
        ops.upsample2x(xin, y)
        return nhwc_to_nchw_f32(y)


class DFL(nn.Module):
    """Holds the frozen 0..15 projection (state_dict key ``head.dfl.conv.weight``,
    components.py:162-174).  The softmax-expectation itself is fused into the head-decode kernel."""

    def __init__(self, ch=16):
        super().__init__()
        self.ch = ch
        self.conv = nn.Conv2d(ch, 1, kernel_size=1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(ch, dtype=torch.float).view(1, ch, 1, 1)

    def forward(self, x):
        raise YmsError("DFL runs fused inside yms_head_decode; call Head(...) in eval mode")


def _slot(block: str, cin: int, cout: int, n: int, k: int):
    if block == "c2f":
        return C2f(cin, cout, num_bottlenecks=n)
    if block == "ms":
        return MSBlock(cin, cout, kernel_size=k, layers_num=n)
    raise ValueError(f"unknown block type {block!r} (expected 'c2f' or 'ms')")


# =============================================================================================
# backbone / neck / head / model
# =============================================================================================
class Backbone(_Compiled):
    def __init__(self, version, in_channels=3, shortcut=True, block="c2f"):
        super().__init__()
        d, w, r = yolo_params(version)
        c1, c2, c3, c4, c5 = int(64 * w), int(128 * w), int(256 * w), int(512 * w), int(512 * w * r)
        self.conv0 = Conv(in_channels, c1, kernel_size=3, stride=2, padding=1)
        self.conv1 = Conv(c1, c2, kernel_size=3, stride=2, padding=1)
        self.conv3 = Conv(c2, c3, kernel_size=3, stride=2, padding=1)
        self.conv5 = Conv(c3, c4, kernel_size=3, stride=2, padding=1)
        self.conv7 = Conv(c4, c5, kernel_size=3, stride=2, padding=1)
        self.c2f_2 = _slot(block, c2, c2, int(3 * d), 3)
        self.c2f_4 = _slot(block, c3, c3, int(6 * d), 3)
        self.c2f_6 = _slot(block, c4, c4, int(6 * d), 5)
        self.c2f_8 = _slot(block, c5, c5, int(3 * d), 7)
        self.sppf = SPPF(c5, c5, kernel_size=5)
        self.out_channels = (c3, c4, c5)

    def emit(self, P, image, outs=(None, None, None)):
        """image: ImageSlot holding the caller's f32 NCHW tensor (read in place by the stem kernel,
        which therefore stays outside the CUDA graph).  outs: optional destination views."""
        b, cin, h, w = image.shape
        if cin != 3:
            raise YmsError("Backbone: the stem kernel expects 3 input channels")
        if h % 32 or w % 32:
            raise YmsError("input height and width must be multiples of 32")   # reference: torch.cat raises
        w0, b0 = self.conv0.folded()
        w0, b0 = w0.contiguous(), b0.contiguous()
        y0 = P.buf(b, h // 2, w // 2, self.conv0.conv.out_channels)
        P.hold(w0, b0)
        assert not P.steps, "the stem must be the first step of the program"
        if image.u8:     # raw uint8 HWC images: ToTensor + Normalize fused into the stem
            P.add(lambda: ops.stem_conv_u8(image.tensor, w0, b0, y0), nbytes=1.0 * b * cin * h * w + 2.0 * y0.numel(),
                  flops=2.0 * y0.numel() * 27, name=f"stem(u8) 3->{y0.shape[3]} @{h}x{w}")
        else:
            P.add(lambda: ops.stem_conv(image.tensor, w0, b0, y0), nbytes=4.0 * b * cin * h * w + 2.0 * y0.numel(),
                  flops=2.0 * y0.numel() * 27, name=f"stem 3->{y0.shape[3]} @{h}x{w}")
        P.eager_prefix = 1
        x = self.conv1.emit(P, y0)
        x = self.c2f_2.emit(P, x)
        x = self.conv3.emit(P, x)
        p3 = self.c2f_4.emit(P, x, out=outs[0])
        x = self.conv5.emit(P, p3)
        p4 = self.c2f_6.emit(P, x, out=outs[1])
        x = self.conv7.emit(P, p4)
        x = self.c2f_8.emit(P, x)
        p5 = self.sppf.emit(P, x, out=outs[2])
        return p3, p4, p5

    def forward(self, x):
        prog, io = _get_program(self, (x,), lambda P, xs: self.emit(P, xs[0]), image_input=True)
        io["inputs"][0].bind(x)
        prog.run()
        return tuple(nhwc_to_nchw_f32(t) for t in io["outputs"])


class Neck(_Compiled):
    def __init__(self, version, block="c2f"):
        super().__init__()
        d, w, r = yolo_params(version)
        c3, c4, c5 = int(256 * w), int(512 * w), int(512 * w * r)
        n = int(3 * d)
        self.up = Upsample()
        self.c2f_1 = _slot(block, int(512 * w * (1 + r)), c4, n, 5)
        self.c2f_2 = _slot(block, int(768 * w), c3, n, 3)
        self.c2f_3 = _slot(block, int(768 * w), c4, n, 5)
        self.c2f_4 = _slot(block, int(512 * w * (1 + r)), c5, n, 7)
        self.conv1 = Conv(c3, c3, kernel_size=3, stride=2, padding=1)
        self.conv2 = Conv(c4, c4, kernel_size=3, stride=2, padding=1)
        self.channels = (c3, c4, c5)

    def alloc(self, P, b, h3, w3):
        """Concat buffers of yolov8_neck.py:76-92; returns the slices the backbone must fill."""
        c3, c4, c5 = self.channels
        h4, w4, h5, w5 = h3 // 2, w3 // 2, h3 // 4, w3 // 4
        cat1 = P.buf(b, h4, w4, c5 + c4)        # [up(P5) | P4]
        cat2 = P.buf(b, h3, w3, c4 + c3)        # [up(res2) | P3]
        cat3 = P.buf(b, h4, w4, c3 + c4)        # [conv1(out1) | res2]
        cat4 = P.buf(b, h5, w5, c4 + c5)        # [conv2(out2) | P5]
        self.__dict__["_cats"] = (cat1, cat2, cat3, cat4)
        return cat2[..., c4:], cat1[..., c5:], cat4[..., c4:]

    def emit(self, P, on_output=None):
        """on_output(i, tensor): called as soon as output i (N3, N4, N5 in this order) has been emitted -- the model starts the
        head of that scale there, on its own graph branch, while the rest of the neck runs."""
        c3, c4, c5 = self.channels
        cat1, cat2, cat3, cat4 = self.__dict__.pop("_cats")
        on_output = on_output or (lambda i, t: None)
        # C2f / MS-Block consumers take the upsample + concat inside their first 1x1 conv (emit(..., up_src)); other blocks read the
        # concat buffer that upsample2x_kernel fills
        def foldable(block):              # limits of yms_conv_plan_add_upsampled
            if not FUSE_UPSAMPLE or not isinstance(block, (C2f, MSBlock)):
                return False
            first = block.conv1 if isinstance(block, C2f) else block.in_conv
            c_out = first.conv.out_channels
            return first.conv.groups == 1 and c_out % (16 if c_out <= 256 else 64) == 0
        if foldable(self.c2f_1):
            res2 = self.c2f_1.emit(P, cat1, out=cat3[..., c3:], up_src=cat4[..., c4:])
        else:
            self.up.emit(P, cat4[..., c4:], cat1[..., :c5])
            res2 = self.c2f_1.emit(P, cat1, out=cat3[..., c3:])
        if foldable(self.c2f_2):
            out1 = self.c2f_2.emit(P, cat2, up_src=res2)
        else:
            self.up.emit(P, res2, cat2[..., :c4])
            out1 = self.c2f_2.emit(P, cat2)
        on_output(0, out1)
        self.conv1.emit(P, out1, out=cat3[..., :c3])
        out2 = self.c2f_3.emit(P, cat3)
        on_output(1, out2)
        self.conv2.emit(P, out2, out=cat4[..., :c4])
        out3 = self.c2f_4.emit(P, cat4)
        on_output(2, out3)
        return out1, out2, out3

    def forward(self, x_res_1, x_res_2, x):
        def build(P, xs):
            b, h3, w3, _ = xs[0].shape
            dsts = self.alloc(P, b, h3, w3)
            for d, s in zip(dsts, xs):
                P.add((lambda d=d, s=s: d.copy_(s)))      # API adaptation for the stand-alone neck only
            return self.emit(P)
        return _run_standalone(self, (x_res_1, x_res_2, x), build)


class Head(_Compiled):
    def __init__(self, version, ch=16, num_classes=80):
        super().__init__()
        self.ch = ch
        self.coordinates = self.ch * 4
        self.num_classes = num_classes
        self.no = self.coordinates + num_classes
        self.stride = torch.zeros(3)                     # plain attribute, like the reference (:79)
        d, w, r = yolo_params(version)
        chans = (int(256 * w), int(512 * w), int(512 * w * r))

        def branch(cin, cout):
            return nn.Sequential(Conv(cin, cout, kernel_size=3, stride=1, padding=1),
                                 Conv(cout, cout, kernel_size=3, stride=1, padding=1),
                                 nn.Conv2d(cout, cout, kernel_size=1, stride=1))

        self.box = nn.ModuleList([branch(c, self.coordinates) for c in chans])
        self.cls = nn.ModuleList([branch(c, num_classes) for c in chans])
        self.dfl = DFL()                                  # the reference ignores `ch` here too (:113)

    @property
    def nc_pad(self) -> int:
        """Class channels as the kernels see them: the class branch is zero-padded to a multiple of 16 when the weights are
        packed (state_dict keys and shapes stay the reference's).  Padded channels carry logit -1e4 -> score exactly 0, so
        they never win the arg-max; predictions and raw outputs are sliced back to num_classes."""
        return (self.num_classes + 15) // 16 * 16

    def can_fuse_decode(self) -> bool:
        """The decode-fused epilogue handles (padded) class counts up to 128 (include/yms_b200.h)."""
        return FUSE_DECODE and self.ch == 16 and 0 < self.nc_pad <= 128

    def _cls_padded(self, i):
        """Folded weights of cls[i] with every class dimension zero-padded from num_classes to nc_pad."""
        nc, ncp = self.num_classes, self.nc_pad
        w0, b0 = self.cls[i][0].folded()
        w1, b1 = self.cls[i][1].folded()
        w2, b2 = self.cls[i][2].weight.detach().float(), self.cls[i][2].bias.detach().float()
        if ncp != nc:
            pad_o = lambda t: torch.cat([t, t.new_zeros((ncp - nc,) + tuple(t.shape[1:]))], 0)
            pad_i = lambda t: torch.cat([t, t.new_zeros((t.shape[0], ncp - nc) + tuple(t.shape[2:]))], 1)
            w0, b0 = pad_o(w0), pad_o(b0)
            w1, b1 = pad_i(pad_o(w1)), pad_o(b1)
            w2 = pad_i(pad_o(w2))
            b2 = torch.cat([b2, b2.new_full((ncp - nc,), -1.0e4)])
        return (w0, b0), (w1, b1), (w2, b2)

    def begin(self, P, shapes, device, fuse_decode: bool = False):
        """Set up one head emission over feature maps of the given [(B, H, W)] shapes; returns the context for emit_scale.

        fuse_decode: the program's final 1x1 convs decode in their epilogue (pred [B,A,4+nc] + candidates, `P.decoded`)
        instead of storing the logits; the logit-storing plans are kept in `P.raw_tail` and only run when the raw
        tensors are asked for (training-mode output, forward_raw)."""
        if self.ch != 16:
            raise RuntimeError("DFL is fixed to 16 bins (the reference builds DFL() with its default ch)")
        ctx = {"fuse": fuse_decode, "bases": [], "dec": None}
        base = 0
        for (_, h, w) in shapes:
            ctx["bases"].append(base)
            base += h * w
        if fuse_decode:
            b = shapes[0][0]
            dec = {"pred": torch.empty((b, base, 4 + self.nc_pad), dtype=torch.float32, device=device),
                   "boxes": torch.empty((b, base, 4), dtype=torch.float32, device=device),
                   "scores": torch.empty((b, base), dtype=torch.float32, device=device),
                   "labels": torch.empty((b, base), dtype=torch.int32, device=device),
                   "stride": torch.zeros(4, dtype=torch.float32, device=device), "stride_vals": None}
            P.hold(*[v for v in dec.values() if torch.is_tensor(v)])
            P.decoded = dec
            ctx["dec"] = dec
        return ctx

    def emit_scale(self, P, ctx, i: int, f: torch.Tensor) -> torch.Tensor:
        """Head branches of scale i over feature map f -> the fp32 raw tensor [B,H,W,64+nc_pad] (box | cls),
        yolov8_head.py:119-122 for one i."""
        ncp = self.nc_pad
        nop = self.coordinates + ncp
        dec, fuse_decode = ctx["dec"], ctx["fuse"]
        b, h, w, _ = f.shape
        raw = P.buf(b, h, w, nop, dtype=torch.float32)
        # box[i][0] and cls[i][0] read the same feature map: ONE 3x3 conv with the two weight sets
        # stacked along c_out (64 + nc) reads it once; the second convs take channel slices.
        wb, bb = self.box[i][0].folded()
        (wc, bc), (wc1, bc1), (wc2, bc2) = self._cls_padded(i)
        first = P.buf(b, h, w, nop)
        P.conv(pack_weight(torch.cat([wb, wc], 0)), torch.cat([bb, bc]).contiguous(), f, first, ksize=3, stride=1, act=True)
        for kind, lo, hi in (("box", 0, self.coordinates), ("cls", self.coordinates, nop)):
            if kind == "box":
                t = self.box[i][1].emit(P, first[..., lo:hi])
                last = self.box[i][2]
                wl, bl = pack_weight(last.weight.detach().float()), last.bias.detach().float().contiguous()
            else:
                t = P.buf(b, h, w, ncp)
                P.conv(pack_weight(wc1), bc1.contiguous(), first[..., lo:hi], t, ksize=3, stride=1, act=self.cls[i][1].has_act)
                wl, bl = pack_weight(wc2), bc2.contiguous()
            if fuse_decode:
                cand = {"cand_boxes": dec["boxes"]} if lo == 0 else {"cand_scores": dec["scores"], "cand_labels": dec["labels"]}
                P.conv(wl, bl, t, raw[..., lo:hi], ksize=1, stride=1, act=False,
                       decode=dict(branch=kind, stride=dec["stride"][i:i + 1], pred=dec["pred"], anchor_base=ctx["bases"][i], **cand))
                P.raw_tail.append(P.conv(wl, bl, t, raw[..., lo:hi], ksize=1, stride=1, act=False, scheduled=False))
            else:
                P.conv(wl, bl, t, raw[..., lo:hi], ksize=1, stride=1, act=False)
        return raw

    def emit(self, P, feats: Sequence[torch.Tensor], fuse_decode: bool = False) -> List[torch.Tensor]:
        """-> 3 fp32 raw tensors [B,H,W,64+nc_pad] (box | cls), yolov8_head.py:119-122: the three scales as parallel branches."""
        ctx = self.begin(P, [(f.shape[0], f.shape[1], f.shape[2]) for f in feats], feats[0].device, fuse_decode)
        raws = []
        P.fork(len(feats))                                   # the scales are independent: parallel graph branches
        for i, f in enumerate(feats):
            P.branch(i)
            raws.append(self.emit_scale(P, ctx, i, f))
        P.join()
        return raws

    def _stride_list(self):
        st = self.stride
        key = (id(st), getattr(st, "_version", 0))
        cached = self.__dict__.get("_stride_cache")
        if cached is None or cached[0] != key:
            vals = [float(v) for v in (st.tolist() if torch.is_tensor(st) else st)]
            cached = (key, vals)
            self.__dict__["_stride_cache"] = cached
        return cached[1]

    def decode(self, raws):
        """raws: the program's [B,H,W,64+nc_pad] logits -> [B,A,4+num_classes] (a view when classes were padded)."""
        return ops.head_decode(raws, self._stride_list(), self.nc_pad)[..., :4 + self.num_classes]

    def forward(self, x):
        raws = _run_standalone(self, tuple(x), lambda P, xs: tuple(self.emit(P, xs)), raw_out=True)
        if self.training:
            return [r.permute(0, 3, 1, 2)[:, :self.no] for r in raws]
        return self.decode([r.contiguous() for r in raws])


class YOLOv8(_Compiled):
    """yolov8/yolov8.py:8-31.  eval: [B, A, 4+nc] fp32; train: list of 3 raw [B, 64+nc, H, W]."""

    def __init__(self, version: str, num_classes: int, dfl_ch: int = 16, block: str = "c2f"):
        super().__init__()
        self.backbone = Backbone(version, block=block)
        self.neck = Neck(version, block=block)
        self.head = Head(version=version, num_classes=num_classes, ch=dfl_ch)

    def _build(self, P, xs):
        img = xs[0]
        b, _, h, w = img.shape
        outs = self.neck.alloc(P, b, h // 8, w // 8)
        ps = self.backbone.emit(P, img, outs=outs)
        # The head of a scale only needs that scale's neck output: scales 0 and 1 start on their own graph branches as soon as
        # N3 / N4 exist and run next to the rest of the neck (their launches fill the drain / launch gaps of the neck's chain
        # and the SMs its small-map layers leave idle); scale 2 continues the main chain.
        ctx = self.head.begin(P, [(b, h // s, w // s) for s in (8, 16, 32)], img_device(img), fuse_decode=self.head.can_fuse_decode())
        raws = [None, None, None]

        def start_head(i, t):
            if i < 2 and OVERLAP_HEAD:
                P.fork_one(i + 1)
                P.branch(i + 1)
                raws[i] = self.head.emit_scale(P, ctx, i, t)
                P.branch(0)
            else:
                raws[i] = self.head.emit_scale(P, ctx, i, t)

        feats = self.neck.emit(P, start_head)
        P.join()
        self.__dict__["_taps"] = {"p": ps, "n": feats}      # NHWC bf16 views, for the parity tests
        return tuple(raws)

    def _run(self, x):
        """Replay the program on x.  Decode-fused programs leave pred / candidates in `prog.decoded`; the stride values
        live in device memory and follow head.stride (a plain attribute in the reference) from call to call."""
        prog, io = _get_program(self, (x,), self._build, image_input=True)
        io["inputs"][0].bind(x)
        self._sync_stride(prog)
        prog.run()
        return prog, io

    def _sync_stride(self, prog):
        """head.stride is a plain attribute in the reference: the program reads it from device memory, refreshed when it changes."""
        dec = prog.decoded
        if dec is not None:
            vals = self.head._stride_list()
            if dec["stride_vals"] != vals:
                if len(vals) != 3:
                    raise YmsError("head.stride must hold 3 values")
                dec["stride"][:3].copy_(torch.tensor(vals, dtype=torch.float32))
                dec["stride_vals"] = list(vals)

    def forward_raw(self, x):
        """Run the network; returns the program's static fp32 raw head buffers [B,H,W,64+nc_pad] (nc_pad = classes rounded up to 16)."""
        prog, io = self._run(x)
        for plan in prog.raw_tail:        # decode-fused program: the logits are only stored on request
            plan.run()
        return io["outputs"]

    def forward(self, x):
        if self.head.training:
            return [r.permute(0, 3, 1, 2)[:, :self.head.no].clone() for r in self.forward_raw(x)]
        prog, io = self._run(x)
        if prog.decoded is not None:                       # the program's buffer is overwritten by the next call
            return prog.decoded["pred"][..., :4 + self.head.num_classes].clone()
        return self.head.decode(io["outputs"])

    @torch.no_grad()
    def detect(self, x, conf_thresh: float = 0.25, iou_thresh: float = 0.45, max_det: Optional[int] = None):
        """Fused forward + decode + class-aware NMS (the reference's tools/test.py:160-218 per batch).
        Returns (boxes [B,A,4], scores [B,A], labels [B,A], keep [B,A], count [B]) -- plus dets [B,max_det,6] rows
        (x1,y1,x2,y2,score,label; label -1 = unused) when max_det is given.  All of them are the program's STATIC buffers
        (valid until the next call on this model with the same input shape): nothing is allocated per call.

        Steady state: when the same input tensor storage shows up again (a caller recycling its input buffers, as a
        pipelined inference loop does), the WHOLE step -- stem, every conv / glue launch, decode, NMS, gather -- is replayed
        as ONE CUDA graph captured for that input address: one host call per step.  Other calls run the stem and the
        post-process as separate launches around the program's graph."""
        prog, io = _get_program(self, (x,), self._build, image_input=True)
        slot = io["inputs"][0]
        slot.bind(x)
        self._sync_stride(prog)
        nc = self.head.nc_pad
        key = (slot.tensor.data_ptr(), float(conf_thresh), float(iou_thresh), max_det)
        full = prog.full_graphs.get(key)
        if full is not None:
            full[0].replay()
            return full[1]
        b = slot.shape[0]
        anchors = sum(t.shape[1] * t.shape[2] for t in io["outputs"])
        post = prog.post.get(max_det)
        if post is None:
            post = prog.post[max_det] = ops.PostBuffers(b, anchors, prog.device, max_det)

        def step(run_program, pb):
            run_program()
            if prog.decoded is not None:
                boxes, scores, labels = prog.decoded["boxes"], prog.decoded["scores"], prog.decoded["labels"]
            else:
                _, (boxes, scores, labels) = ops.head_decode(io["outputs"], self.head._stride_list(), nc, with_candidates=True)
            keep, count = ops.nms_batched(boxes, scores, labels, conf_thresh, iou_thresh, nc, out=pb)
            outs = (boxes, scores, labels, keep, count)
            if max_det is not None:
                outs = outs + (ops.gather_detections(boxes, scores, labels, keep, count, max_det, out=pb.dets),)
            return outs

        if len(prog.seen) > 64:                               # a caller that never recycles its input buffers
            prog.seen.clear()
        seen = prog.seen.get(key, 0) + 1
        prog.seen[key] = seen
        if seen >= 2 and prog.decoded is not None and len(prog.full_graphs) < 8:
            # every captured input address gets its OWN keep / count / dets buffers: a pipelined caller reads the results of
            # buffer A (device -> host copy on another stream) while the step on buffer B is already running
            pb = ops.PostBuffers(b, anchors, prog.device, max_det)
            torch.cuda.synchronize(prog.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs = step(prog.run_eager, pb)
            prog.full_graphs[key] = (g, outs)
            prog.full_keep.append((slot.tensor, pb))          # the captured addresses must stay alive
            g.replay()
            return outs
        return step(prog.run, post)


# =============================================================================================
# program cache
# =============================================================================================
class ImageSlot:
    """The caller's image batch, rebound on every call (no staging copy): NCHW fp32 (the reference's
    interface) or raw uint8 NHWC [B,H,W,3] (normalisation fused into the stem)."""

    def __init__(self, x: torch.Tensor):
        self.u8 = x.dtype == torch.uint8
        if self.u8:
            if x.dim() != 4 or x.shape[3] != 3:
                raise YmsError("uint8 images must be NHWC [B,H,W,3]")
            self.shape = (x.shape[0], 3, x.shape[1], x.shape[2])     # logical NCHW shape
        else:
            self.shape = tuple(x.shape)
        self.tensor = None
        self.bind(x)

    def bind(self, x: torch.Tensor):
        if self.u8:
            if x.dtype != torch.uint8:
                raise YmsError("this program was compiled for uint8 images")
            x = x.contiguous()
        elif x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        self.tensor = x

def _get_program(module: _Compiled, xs: Tuple[torch.Tensor, ...], build, image_input=False):
    for x in xs:
        if not torch.is_tensor(x) or not x.is_cuda:
            raise YmsError("yolo_ms_b200 modules run on CUDA tensors only (no CPU fallback)")
    dev = xs[0].device
    if _device_of(module) != dev:
        raise YmsError("module parameters and input live on different devices")
    key = tuple((tuple(x.shape), str(x.dtype)) for x in xs) + (str(dev),)
    cache = module._programs()
    hit = cache.get(key)
    if hit is not None:
        return hit
    with torch.no_grad():
        P = Program(dev)
        if image_input:
            ins = [ImageSlot(x) for x in xs]
        else:
            ins = [torch.empty((x.shape[0], x.shape[2], x.shape[3], x.shape[1]), dtype=torch.bfloat16, device=dev) for x in xs]
        if not image_input:
            P.hold(*ins)
        outs = build(P, ins)
        P.capture()
    hit = (P, {"inputs": ins, "outputs": list(outs)})
    cache[key] = hit
    return hit


def _run_standalone(module, xs, build, raw_out=False):
    """Stand-alone sub-module call with the reference's NCHW fp32 tensors at the boundary."""
    prog, io = _get_program(module, tuple(xs), build)
    for dst, x in zip(io["inputs"], xs):
        dst.copy_(x.permute(0, 2, 3, 1))
    prog.run()
    if raw_out:
        return io["outputs"]
    return tuple(nhwc_to_nchw_f32(t) for t in io["outputs"])
