"""Functional fp32 CPU restatement of the reference model forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Operates on a plain state_dict with the
reference's key names; every function cites the reference lines it follows
(paths relative to /root/reference).  Uses the same ATen CPU ops the reference reaches
(conv2d / batch_norm / silu / max_pool2d / interpolate / softmax / sigmoid), so it also
serves as the CPU-baseline timing port in bench.py.

``block='ms'`` swaps every C2f slot for the repo-local MS-Block (``ms_block``); that
variant is PARITY-UNPINNED -- the reference has no MS-Block.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # yolov8/model/components.py:73


def width_params(version: str):
    """yolov8/model/components.py:193-209 (depth, width, ratio)."""
    table = {"n": (1 / 3, 1 / 4, 2.0), "s": (1 / 3, 1 / 2, 2.0), "m": (2 / 3, 3 / 4, 1.5),
             "l": (1.0, 1.0, 1.0), "x": (1.0, 1.25, 1.0)}
    if version not in table:
        raise ValueError(f"Unknown YOLOv8 version: {version}")
    return table[version]


_CALIBRATING = False
# When True the oracle keeps the reference's algorithm but applies the PRODUCT'S numeric contract
# (DESIGN.md): BN folded into the weights in fp32, weights and every stored activation rounded to
# bf16, fp32 accumulation, fp32 head logits / decode.  The CUDA path must agree with this variant
# to ~1e-2 END TO END, which makes wiring bugs visible that the bf16-vs-fp32 noise floor would hide.
EMULATE_BF16 = False


def _bf16(t):
    return t.to(torch.bfloat16).float()


def conv_unit(sd, p, x, stride=1, act=True, residual=None, first=False):
    """Conv2d(bias=False) -> BN(eval, eps 1e-3) -> SiLU (+ residual).  components.py:69-77, :91-92."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    groups = x.shape[1] // w.shape[1]
    if EMULATE_BF16 and not _CALIBRATING:
        scale = sd[p + ".bn.weight"] / torch.sqrt(sd[p + ".bn.running_var"] + BN_EPS)
        wf = w * scale.view(-1, 1, 1, 1)
        bf = sd[p + ".bn.bias"] - sd[p + ".bn.running_mean"] * scale
        if groups == 1:                     # tensor-core kernels: bf16 operands (the depthwise kernel keeps fp32 weights)
            wf = _bf16(wf)
        if first:                           # the stem kernel converts the fp32 image to bf16 while gathering
            x = _bf16(x)
        y = F.conv2d(x, wf, bf, stride, k // 2, 1, groups)
        y = F.silu(y) if act else y
        if residual is not None:
            y = y + residual
        return _bf16(y)
    y = F.conv2d(x, w, None, stride, k // 2, 1, groups)
    if _CALIBRATING:  # test-weight generation only: set running stats from this batch
        yd = y.double()
        sd[p + ".bn.running_mean"] = yd.mean((0, 2, 3)).float()
        sd[p + ".bn.running_var"] = yd.var((0, 2, 3), unbiased=False).clamp_min(1e-6).float()
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"],
                     sd[p + ".bn.weight"], sd[p + ".bn.bias"], False, 0.0, BN_EPS)
    y = F.silu(y) if act else y
    return y if residual is None else y + residual


def _count(sd, prefix):
    n = 0
    while f"{prefix}.{n}.conv1.conv.weight" in sd:
        n += 1
    return n


def c2f(sd, p, x):
    """components.py:108-122: 1x1, split halves, chain the FIRST half through the
    bottlenecks, concat order [b_n, ..., b_1, x1, x2], 1x1.  Bottleneck (components.py:87-93)
    always adds its input (C2f never forwards ``shortcut``, components.py:104)."""
    y = conv_unit(sd, p + ".conv1", x)
    half = y.shape[1] // 2
    x1, x2 = y[:, :half], y[:, half:]
    outs = [x1, x2]
    for j in range(_count(sd, p + ".m")):
        t = conv_unit(sd, f"{p}.m.{j}.conv1", x1)
        x1 = conv_unit(sd, f"{p}.m.{j}.conv2", t, residual=x1)      # x += x_in (components.py:91-92)
        outs.insert(0, x1)
    return conv_unit(sd, p + ".conv2", torch.cat(outs, 1))


def ms_block(sd, p, x):
    """Repo-local MS-Block (after arXiv 2308.05480; NOT in the reference -> unpinned).

    in_conv 1x1 (C_in -> 3*c); split in 3 branches of c channels; branch 0 identity;
    branch i>=1: input x_i + y_{i-1}, then L x [1x1 (c->2c), depthwise kxk (2c), 1x1 (2c->c)];
    concat the 3 branch outputs; out_conv 1x1 (3c -> C_out).  Every sub-layer is a
    reference-style Conv triplet (conv+BN+SiLU, components.py:69-77)."""
    y = conv_unit(sd, p + ".in_conv", x)
    c = y.shape[1] // 3
    outs = [y[:, :c]]
    for b in (1, 2):
        t = y[:, b * c:(b + 1) * c] + outs[-1]
        layer = 0
        while f"{p}.branches.{b - 1}.{layer}.pw1.conv.weight" in sd:
            q = f"{p}.branches.{b - 1}.{layer}"
            t = conv_unit(sd, q + ".pw1", t)
            t = conv_unit(sd, q + ".dw", t)
            t = conv_unit(sd, q + ".pw2", t)
            layer += 1
        outs.append(t)
    return conv_unit(sd, p + ".out_conv", torch.cat(outs, 1))


def csp_slot(sd, p, x):
    return ms_block(sd, p, x) if (p + ".in_conv.conv.weight") in sd else c2f(sd, p, x)


def sppf(sd, p, x):
    """components.py:138-150: 1x1, three chained 5x5/s1/p2 max pools, concat, 1x1."""
    x = conv_unit(sd, p + ".conv1", x)
    x1 = F.max_pool2d(x, 5, 1, 2)
    x2 = F.max_pool2d(x1, 5, 1, 2)
    x3 = F.max_pool2d(x2, 5, 1, 2)
    return conv_unit(sd, p + ".conv2", torch.cat([x, x1, x2, x3], 1))


def backbone(sd, x, taps=None):
    """yolov8_backbone.py:54-74."""
    x = conv_unit(sd, "backbone.conv0", x, 2, first=True)
    if taps is not None:
        taps["backbone.conv0"] = x
    x = conv_unit(sd, "backbone.conv1", x, 2)
    x = csp_slot(sd, "backbone.c2f_2", x)
    x = conv_unit(sd, "backbone.conv3", x, 2)
    p3 = csp_slot(sd, "backbone.c2f_4", x)
    x = conv_unit(sd, "backbone.conv5", p3, 2)
    p4 = csp_slot(sd, "backbone.c2f_6", x)
    x = conv_unit(sd, "backbone.conv7", p4, 2)
    x = csp_slot(sd, "backbone.c2f_8", x)
    p5 = sppf(sd, "backbone.sppf", x)
    return p3, p4, p5


def neck(sd, p3, p4, p5):
    """yolov8_neck.py:67-94 (nearest x2 upsample = components.py:159-160)."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    r2 = csp_slot(sd, "neck.c2f_1", torch.cat([up(p5), p4], 1))
    n3 = csp_slot(sd, "neck.c2f_2", torch.cat([up(r2), p3], 1))
    n4 = csp_slot(sd, "neck.c2f_3", torch.cat([conv_unit(sd, "neck.conv1", n3, 2), r2], 1))
    n5 = csp_slot(sd, "neck.c2f_4", torch.cat([conv_unit(sd, "neck.conv2", n4, 2), p5], 1))
    return n3, n4, n5


def head_raw(sd, feats):
    """yolov8_head.py:119-122: per scale cat(box branch, cls branch) -> [B, 64+nc, H, W]."""
    out = []
    for i, f in enumerate(feats):
        branch = []
        for name in ("box", "cls"):
            t = conv_unit(sd, f"head.{name}.{i}.0", f)
            t = conv_unit(sd, f"head.{name}.{i}.1", t)
            wl = sd[f"head.{name}.{i}.2.weight"]
            t = F.conv2d(t, _bf16(wl) if EMULATE_BF16 else wl, sd[f"head.{name}.{i}.2.bias"])
            branch.append(t)
        out.append(torch.cat(branch, 1))
    return out


def decode(raw, strides, reg_max=16):
    """yolov8_head.py:127-144 + make_anchors :146-158 + DFL components.py:176-191.

    raw: list of [B, 4*reg_max+nc, H_i, W_i]; strides: 3 floats.  -> [B, A, 4+nc]."""
    b = raw[0].shape[0]
    no = raw[0].shape[1]
    anchors, svec = [], []
    for r, s in zip(raw, strides):
        h, w = r.shape[2:]
        sx = torch.arange(w, dtype=r.dtype, device=r.device) + 0.5
        sy = torch.arange(h, dtype=r.dtype, device=r.device) + 0.5
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        anchors.append(torch.stack((gx, gy), -1).view(-1, 2))
        svec.append(torch.full((h * w, 1), float(s), dtype=r.dtype, device=r.device))
    anchors = torch.cat(anchors).t()           # [2, A]
    svec = torch.cat(svec).t()                 # [1, A]
    x = torch.cat([r.reshape(b, no, -1) for r in raw], 2)
    box, cls = x.split((4 * reg_max, no - 4 * reg_max), 1)
    a = box.shape[2]
    prob = box.view(b, 4, reg_max, a).transpose(1, 2).softmax(1)
    proj = torch.arange(reg_max, dtype=x.dtype, device=x.device).view(1, reg_max, 1, 1)
    dist = F.conv2d(prob, proj.view(1, reg_max, 1, 1)).view(b, 4, a)
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    boxes = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1)
    out = torch.cat((boxes * svec, cls.sigmoid()), 1)
    return out.transpose(1, 2).contiguous()


def calibrate_bn(sd, x):
    """Overwrite every BN running_mean/var in ``sd`` with the statistics of one forward over
    ``x`` (weight generation for tests; see oracle/weights.py)."""
    global _CALIBRATING
    _CALIBRATING = True
    try:
        forward(sd, x)
    finally:
        _CALIBRATING = False
    return sd


def forward_bf16_contract(sd, x, strides=(8.0, 16.0, 32.0), return_parts=False):
    """Same algorithm under the product's numeric contract (see EMULATE_BF16)."""
    global EMULATE_BF16
    EMULATE_BF16 = True
    try:
        return forward(sd, x, strides, return_parts)
    finally:
        EMULATE_BF16 = False


def forward(sd, x, strides=(8.0, 16.0, 32.0), return_parts=False):
    """yolov8/yolov8.py:23-31, eval mode."""
    p3, p4, p5 = backbone(sd, x)
    n3, n4, n5 = neck(sd, p3, p4, p5)
    raw = head_raw(sd, [n3, n4, n5])
    pred = decode(raw, strides)
    if return_parts:
        return {"p": (p3, p4, p5), "n": (n3, n4, n5), "raw": raw, "pred": pred}
    return pred


# ---------------------------------------------------------------------------
# manifest of the MS variant (what the product's block='ms' modules must expose)
# ---------------------------------------------------------------------------
MS_KERNELS = {"backbone.c2f_2": 3, "backbone.c2f_4": 3, "backbone.c2f_6": 5, "backbone.c2f_8": 7,
              "neck.c2f_1": 5, "neck.c2f_2": 3, "neck.c2f_3": 5, "neck.c2f_4": 7}


def ms_manifest_from_c2f(manifest):
    """Derive the (key -> shape) manifest of the block='ms' model from a reference (C2f)
    manifest: same slots, same (in, out) channels, layers = the slot's bottleneck count."""
    from collections import OrderedDict
    out = OrderedDict()
    done = set()

    def unit(p, cout, cin_per_group, k):
        out[p + ".conv.weight"] = [cout, cin_per_group, k, k]
        for s in ("weight", "bias", "running_mean", "running_var"):
            out[f"{p}.bn.{s}"] = [cout]
        out[p + ".bn.num_batches_tracked"] = []

    for key, shape in manifest.items():
        slot = next((s for s in MS_KERNELS if key.startswith(s + ".")), None)
        if slot is None:
            out[key] = shape
            continue
        if slot in done:
            continue
        done.add(slot)
        cin = manifest[slot + ".conv1.conv.weight"][1]
        cout = manifest[slot + ".conv2.conv.weight"][0]
        layers = 0
        while f"{slot}.m.{layers}.conv1.conv.weight" in manifest:
            layers += 1
        c = cout // 2
        k = MS_KERNELS[slot]
        unit(slot + ".in_conv", 3 * c, cin, 1)
        for b in range(2):
            for l in range(layers):
                q = f"{slot}.branches.{b}.{l}"
                unit(q + ".pw1", 2 * c, c, 1)
                unit(q + ".dw", 2 * c, 1, k)
                unit(q + ".pw2", c, 2 * c, 1)
        unit(slot + ".out_conv", cout, 3 * c, 1)
    return out
