"""GPU pre-processing: the reference's ``T.Resize -> T.ToTensor -> T.Normalize`` chain
(yolov8/tools/test.py:114-119) for raw uint8 RGB images.

* resize: Pillow's fixed-point separable BILINEAR resampling (what torchvision's Resize runs on PIL images), two passes
  of ``yms_resample_u8`` with coefficients computed here exactly as Pillow's ``precompute_coeffs`` /
  ``normalize_coeffs_8bpc`` (src/libImaging/Resample.c) -- bit-exact against Pillow;
* ToTensor + Normalize: fused into the stem kernel (``yms_stem_conv_u8``), so the resized uint8 HWC batch is what the
  model consumes: ``model(preprocess_batch(images, (h, w)))``.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import YmsError, check

PRECISION_BITS = 32 - 8 - 2
_coeff_cache: Dict[Tuple[int, int, str], Tuple[int, torch.Tensor, torch.Tensor]] = {}


def bilinear_coeffs(in_size: int, out_size: int):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter over the whole axis.
    -> (ksize, bounds int32 [out, 2] (first index, count), coeffs int32 [out, ksize])."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = filterscale                                   # filter support 1.0, stretched when down-scaling
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        ww = 0.0
        w = [0.0] * n
        for x in range(n):
            v = (x + lo - center + 0.5) * inv
            v = -v if v < 0.0 else v
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(n):
            kk[xx, x] = w[x] / ww if ww != 0.0 else w[x]
        bounds[xx] = (lo, n)
    fixed = np.trunc(np.where(kk < 0, -0.5, 0.5) + kk * float(1 << PRECISION_BITS)).astype(np.int32)
    return ksize, bounds, fixed


def _device_coeffs(in_size: int, out_size: int, device: torch.device):
    key = (in_size, out_size, str(device))
    hit = _coeff_cache.get(key)
    if hit is None:
        ksize, bounds, coeffs = bilinear_coeffs(in_size, out_size)
        hit = (ksize, torch.from_numpy(bounds).to(device), torch.from_numpy(coeffs).to(device))
        _coeff_cache[key] = hit
    return hit


def resize_u8(img: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """img uint8 [H, W, C] (CUDA, contiguous) -> out uint8 [h, w, C] (CUDA, rows may be a view into a batch buffer);
    == np.asarray(PIL.Image.fromarray(img).resize((w, h), BILINEAR)) bit for bit."""
    if not (img.is_cuda and out.is_cuda):
        raise YmsError("yolo_ms_b200 pre-processing runs on CUDA tensors only (no CPU fallback)")
    if img.dtype != torch.uint8 or out.dtype != torch.uint8 or img.dim() != 3 or out.dim() != 3 or img.shape[2] != out.shape[2]:
        raise YmsError("resize_u8: uint8 [H,W,C] tensors with equal channel count expected")
    if img.stride(2) != 1 or img.stride(1) != img.shape[2] or out.stride(2) != 1 or out.stride(1) != out.shape[2]:
        raise YmsError("resize_u8: pixels must be contiguous (row stride may be arbitrary)")
    h0, w0, c = img.shape
    h1, w1, _ = out.shape
    lib = _lib.load()
    stream = torch.cuda.current_stream(img.device).cuda_stream
    if (h0, w0) == (h1, w1):
        out.copy_(img)
        return out
    cur, cur_h, cur_w = img, h0, w0
    with torch.cuda.device(img.device):
        return _resize_passes(lib, stream, img, out, cur, cur_h, cur_w, h0, w0, h1, w1, c)


def _resize_passes(lib, stream, img, out, cur, cur_h, cur_w, h0, w0, h1, w1, c):
    if w0 != w1:
        ks, b, k = _device_coeffs(w0, w1, img.device)
        dst = out if h0 == h1 else torch.empty((h0, w1, c), dtype=torch.uint8, device=img.device)
        check(lib.yms_resample_u8(cur.data_ptr(), cur_h, cur_w, c, cur.stride(0), dst.data_ptr(), cur_h, w1, dst.stride(0),
                                  b.data_ptr(), k.data_ptr(), ks, 1, stream), "yms_resample_u8")
        cur, cur_w = dst, w1
    if h0 != h1:
        ks, b, k = _device_coeffs(h0, h1, img.device)
        check(lib.yms_resample_u8(cur.data_ptr(), cur_h, cur_w, c, cur.stride(0), out.data_ptr(), h1, cur_w, out.stride(0),
                                  b.data_ptr(), k.data_ptr(), ks, 0, stream), "yms_resample_u8")
    return out


def preprocess_batch(images: Sequence, size: Tuple[int, int], device="cuda") -> torch.Tensor:
    """images: uint8 RGB HWC arrays / tensors of arbitrary sizes -> uint8 [B, h, w, 3] on `device`, resized like the
    reference's T.Resize(size).  Feed the result to YOLOv8 (ToTensor + Normalize run inside the stem kernel)."""
    h, w = size
    dev = torch.device(device)
    batch = torch.empty((len(images), h, w, 3), dtype=torch.uint8, device=dev)
    for i, im in enumerate(images):
        if not torch.is_tensor(im):
            im = np.ascontiguousarray(im)
            im = im if im.flags.writeable else im.copy()          # PIL hands out read-only views
        t = im if torch.is_tensor(im) else torch.from_numpy(im)
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise YmsError("preprocess_batch: uint8 RGB HWC images expected")
        resize_u8(t.to(dev, non_blocking=True).contiguous(), batch[i])
    return batch
