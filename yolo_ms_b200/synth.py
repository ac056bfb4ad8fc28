"""Synthetic weights and images for benchmarks and smoke tests (there are no checkpoints or
datasets offline).  Pure data generation -- no compute path lives here.

Every tensor is a pure function of (seed, key, shape), so the build container, the GPU box and the
CPU oracle all see bit-identical weights.  Conv weights ~ N(0, 1/fan_in); BN affine random; BN
running statistics come from a committed calibration fixture (``synth_bn/*.npz``, produced by
``python -m oracle.make_golden`` with one pass of the CPU oracle over a seeded batch) so that
activations stay O(1) through the ~60 layers and all classes are active.
"""
from __future__ import annotations

import hashlib
import math
import os
from collections import OrderedDict

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
BN_DIR = os.path.join(_HERE, "synth_bn")
CONV_GAIN = 1.0
FINAL_GAIN = 1.5     # final biased 1x1 convs of the head: logits std ~1.5


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def make_state_dict(manifest, seed: int = 0):
    """manifest: mapping key -> shape (e.g. ``{k: v.shape for k, v in model.state_dict().items()}``)."""
    sd = OrderedDict()
    for key, shape in manifest.items():
        g = _gen(seed, key)
        shape = list(shape)
        if key.endswith("num_batches_tracked"):
            t = torch.zeros(shape, dtype=torch.long)
        elif key == "head.dfl.conv.weight":
            t = torch.arange(shape[1], dtype=torch.float32).view(shape)
        elif key.endswith("bn.weight"):
            t = torch.rand(shape, generator=g) * 0.8 + 0.6
        elif key.endswith("bn.bias"):
            t = torch.randn(shape, generator=g) * 0.2
        elif key.endswith("bn.running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif key.endswith("bn.running_var"):
            t = torch.rand(shape, generator=g) * 0.8 + 0.6
        elif key.endswith(".bias"):            # the biased final 1x1 convs of the head
            t = torch.randn(shape, generator=g) * 0.1
        elif key.endswith(".weight") and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            gain = CONV_GAIN if key.endswith("conv.weight") else FINAL_GAIN
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        else:
            raise KeyError(f"no recipe for {key} {shape}")
        sd[key] = t
    return sd


def make_images(batch: int, height: int, width: int, seed: int = 7) -> torch.Tensor:
    """ImageNet-normalised synthetic RGB batch (the reference's input convention,
    yolov8/tools/test.py:114-119)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.rand(batch, 3, height, width, generator=g)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (x - mean) / std


def bn_fixture_path(version: str, block: str, seed: int) -> str:
    return os.path.join(BN_DIR, f"bn_{version}_{block}_seed{seed}.npz")


def synthetic_state_dict(model: torch.nn.Module, version: str, block: str = "c2f", seed: int = 1):
    """State dict for ``model`` from the seeded recipe + the committed BN calibration fixture."""
    manifest = OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())
    sd = make_state_dict(manifest, seed=seed)
    path = bn_fixture_path(version, block, seed)
    if os.path.exists(path):
        with np.load(path) as f:
            for k in f.files:
                if k in sd:
                    sd[k] = torch.from_numpy(f[k].astype(np.float32))
    return sd
