"""Checkpoint ingestion and config loading, mirroring yolov8/tools/utils.py:5-9,45-82 and the
loader of yolov8/tools/test.py:94-105 of the reference."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import yaml


def load_config(config_path):
    """Loads a YAML configuration file (yolov8/tools/utils.py:5-9)."""
    with open(config_path, "r") as f:
        return yaml.safe_load(f)


def extract_state_dict(checkpoint):
    """Every checkpoint layout the reference accepts: a bare state_dict, ``{'model': sd}``, ``{'state_dict': sd}``
    (tools/utils.py:55-62), with or without the DataParallel ``module.`` prefix (tools/utils.py:65-67, tools/test.py:97-104)."""
    sd = checkpoint
    if isinstance(checkpoint, dict):
        if "model" in checkpoint and isinstance(checkpoint["model"], dict):
            sd = checkpoint["model"]
        elif "state_dict" in checkpoint and isinstance(checkpoint["state_dict"], dict):
            sd = checkpoint["state_dict"]
    if not isinstance(sd, dict):
        raise TypeError(f"checkpoint holds a {type(sd).__name__}, not a state_dict")
    if any(k.startswith("module.") for k in sd.keys()):
        sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
    return sd


def load_checkpoint(model, checkpoint_path, strict=True, map_location="cpu", trusted=False):
    """Load a reference ``.pt`` checkpoint into a yolo_ms_b200 model (same state_dict keys as the reference's modules).
    The compiled launch programs (folded BN, packed bf16 weights) are rebuilt lazily on the next forward.
    Checkpoints are un-pickled with ``weights_only=True`` (what the reference's plain ``torch.load`` does on current torch):
    only tensors and plain containers are accepted; ``trusted=True`` allows arbitrary pickled objects for files you made.
    Returns (missing_keys, unexpected_keys) like ``nn.Module.load_state_dict``."""
    if not os.path.exists(checkpoint_path):
        raise FileNotFoundError(f"Checkpoint file not found: {checkpoint_path}")
    sd = extract_state_dict(torch.load(checkpoint_path, map_location=map_location, weights_only=not trusted))
    res = model.load_state_dict(sd, strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)


def load_pretrained_weights(model, pretrained_path, strict=False):
    """yolov8/tools/utils.py:45-82: tolerant loader used by training scripts (prints instead of raising)."""
    if not pretrained_path or not os.path.exists(pretrained_path):
        print("No pretrained weights found, training from scratch")
        return model
    try:
        missing, unexpected = load_checkpoint(model, pretrained_path, strict=strict)
        if missing:
            print(f"Missing keys: {missing}")
        if unexpected:
            print(f"Unexpected keys: {unexpected}")
        print(f"Successfully loaded pretrained weights from {pretrained_path}")
    except Exception as e:  # noqa: BLE001  (the reference swallows every error here)
        print(f"Error loading pretrained weights: {e}")
        print("Training from scratch")
    return model
