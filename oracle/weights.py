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

import json
import os
from collections import OrderedDict

import torch

from yolo_ms_b200.synth import CONV_GAIN, FINAL_GAIN, make_images, make_state_dict  # noqa: F401  (shared seeded recipe)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_manifest(version: str, block: str = "c2f") -> "OrderedDict[str, list]":
    name = f"manifest_{version}.json" if block == "c2f" else f"manifest_{version}_{block}.json"
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return OrderedDict((k, v) for k, v in json.load(f))


def calibrated_state_dict(version: str, seed: int = 0, block: str = "c2f", calib_hw=(320, 320)):
    """Seeded weights + BN running stats calibrated by one pass of the oracle over a seeded
    synthetic batch (statistics accumulated in float64)."""
    from . import yolov8_oracle as O
    sd = make_state_dict(load_manifest(version, block), seed=seed)
    with torch.no_grad():
        O.calibrate_bn(sd, make_images(2, calib_hw[0], calib_hw[1], seed=1000 + seed))
    return sd
