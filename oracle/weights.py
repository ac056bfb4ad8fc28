"""Deterministic, manifest-driven weights for parity tests and benchmarks.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference builds its parameters with ``nn`` default init, which makes every
activation collapse to ~0 after a few eval-mode BN layers so
parity tests on it are vacuous.  Here every tensor of a state_dict is a pure
function of ``(seed, key, shape)``: no dependence on module construction order or on
a data-dependent BN calibration pass, so the build container and the GPU box produce
bit-identical weights.  The ``(key -> shape)`` manifests under ``tests/golden`` were
dumped from the real reference (``oracle/make_golden.py``), so loading these tensors
into ``/root/reference``'s ``YOLOv8`` with ``load_state_dict(strict=True)`` works.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
from collections import OrderedDict

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# Conv weights are N(0, 1/fan_in); BN running stats are then CALIBRATED on a seeded synthetic
# batch (``calibrated_state_dict``) so every pre-activation is ~N(beta, gamma^2): activations
# stay O(1) through all ~60 layers.  The final biased 1x1 convs of the head
# get FINAL_GAIN so logits have std ~1.5 (all 80 classes win somewhere, scores span 0.05..0.97).
CONV_GAIN = 1.0
FINAL_GAIN = 1.5


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def load_manifest(version: str, block: str = "c2f") -> "OrderedDict[str, list]":
    name = f"manifest_{version}.json" if block == "c2f" else f"manifest_{version}_{block}.json"
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return OrderedDict((k, v) for k, v in json.load(f))


def make_state_dict(manifest, seed: int = 0, num_classes_bias: float = 0.0):
    """Return an OrderedDict[str, Tensor] with the manifest's keys/shapes."""
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
        elif key.endswith(".bias"):  # the biased final 1x1 convs of the head
            t = torch.randn(shape, generator=g) * 0.1 + num_classes_bias
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
    /root/reference/yolov8/tools/test.py:114-119)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.rand(batch, 3, height, width, generator=g)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (x - mean) / std


def calibrated_state_dict(version: str, seed: int = 0, block: str = "c2f", calib_hw=(320, 320)):
    """Seeded weights + BN running stats calibrated by one pass of the oracle over a seeded
    synthetic batch (statistics accumulated in float64)."""
    from . import yolov8_oracle as O
    sd = make_state_dict(load_manifest(version, block), seed=seed)
    with torch.no_grad():
        O.calibrate_bn(sd, make_images(2, calib_hw[0], calib_hw[1], seed=1000 + seed))
    return sd
