"""TEST INFRASTRUCTURE ONLY (never imported by yolo_ms_b200/).

CPU restatement of the resize step of the reference's pre-processing, ``T.Resize((h, w))`` applied to a PIL RGB image
(yolov8/tools/test.py:114-119,142-145).  torchvision hands PIL images to ``Image.resize(size, BILINEAR)``; the
arithmetic therefore lives in a third-party dependency that is not vendored by the reference: Pillow (requirements.txt
lists only torch/torchvision; Pillow comes with torchvision; installed here: see PIL.__version__), file
``src/libImaging/Resample.c``.  Its published algorithm is restated below:

  * separable two-pass convolution, horizontal pass first, each pass rounds to uint8;
  * triangle filter whose support is stretched by the down-scaling factor (anti-aliasing), window
    [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image, weights normalised to sum 1 in double;
  * coefficients converted to 22-bit fixed point ``int(0.5 + w * 2**22)``, accumulation in int32 starting from
    ``1 << 21``, result ``clip8(acc >> 22)``.

Pinned against the installed Pillow itself by tests/test_resize.py (bit-exact on random images, up- and down-scaling).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size: int, out_size: int):
    """-> (ksize, bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize]) for the full-image box."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size          # (double)(in1 - in0) / outSize, box as floats
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                                            # bilinear: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)], np.float64)
        ww = 0.0
        for v in w:                                                        # same summation order as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)), (0.5 + kk * (1 << PRECISION_BITS))).astype(np.int64)
    return ksize, bounds, fixed.astype(np.int32)                          # (int) truncation toward zero == astype for these


def _pass(img: np.ndarray, bounds: np.ndarray, coeffs: np.ndarray, axis: int) -> np.ndarray:
    out_size = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.empty(shape, np.uint8)
    src = img.astype(np.int64)
    for o in range(out_size):
        lo, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        k = coeffs[o, :cnt].astype(np.int64)
        if axis == 1:
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, lo:lo + cnt, :], k, axes=([1], [0]))
            out[:, o, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        else:
            acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[lo:lo + cnt, :, :], k, axes=([0], [0]))
            out[o, :, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img uint8 [H, W, 3] -> uint8 [out_h, out_w, 3], == np.asarray(PIL.Image.fromarray(img).resize((out_w, out_h), BILINEAR))."""
    h, w, _ = img.shape
    if (h, w) == (out_h, out_w):
        return img.copy()
    cur = img
    if w != out_w:
        _, bh, kh = precompute_coeffs(w, out_w)
        cur = _pass(cur, bh, kh, axis=1)
    if h != out_h:
        _, bv, kv = precompute_coeffs(h, out_h)
        cur = _pass(cur, bv, kv, axis=0)
    return cur
