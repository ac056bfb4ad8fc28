"""Helpers shared by the -m gpu tests."""
import numpy as np
import torch

DEV = "cuda"


def rel_l2(got, want):
    got = torch.as_tensor(got).detach().float().cpu()
    want = torch.as_tensor(want).detach().float().cpu()
    return float((got - want).norm() / (want.norm() + 1e-30))


def randomize_bn(module, seed=0):
    """Give every BN non-trivial affine + running stats (default init makes BN ~identity)."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            m.weight.data = torch.rand(n, generator=g) * 0.8 + 0.6
            m.bias.data = torch.randn(n, generator=g) * 0.2
            m.running_mean.data = torch.randn(n, generator=g) * 0.2
            m.running_var.data = torch.rand(n, generator=g) * 0.8 + 0.6
    return module


def bf16_round(x):
    return x.to(torch.bfloat16).float()


def prefixed_state(module, prefix):
    return {f"{prefix}.{k}": v.detach().float().cpu() for k, v in module.state_dict().items()}


def iou_xyxy(a, b):
    x1 = np.maximum(a[:, 0], b[:, 0]); y1 = np.maximum(a[:, 1], b[:, 1])
    x2 = np.minimum(a[:, 2], b[:, 2]); y2 = np.minimum(a[:, 3], b[:, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    ua = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter
    return inter / np.maximum(ua, 1e-12)
