"""-m gpu: the depthwise / fused MS-Block layer kernel (csrc/ms_fused.cu) through the C ABI.

Pins: tests/golden/dwconv_ref.npz comes from the REFERENCE's own Conv class (components.py:69-77) instantiated with
groups=c and chained as pw1 -> dw -> pw2 (oracle/make_golden.py::dump_dwconv).  Tolerances: one bf16 output rounding for the
depthwise unit (rel-L2 <= 4e-3); three chained bf16 roundings for the layer (<= 1e-2); the kernel's three fusion modes against
each other <= 3e-3 (same roundings, different fp32 summation order -> isolated bf16 flips)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from gpu_util import DEV, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from yolo_ms_b200 import ops as _ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def _fold(g, prefix):
    """(weight, bias) of a reference Conv unit with BN folded, fp32 (components.py:69-77, eps 1e-3)."""
    w = torch.from_numpy(g[prefix + "conv.weight"])
    scale = torch.from_numpy(g[prefix + "bn.weight"]) / torch.sqrt(torch.from_numpy(g[prefix + "bn.running_var"]) + 1e-3)
    return w * scale.view(-1, 1, 1, 1), torch.from_numpy(g[prefix + "bn.bias"]) - torch.from_numpy(g[prefix + "bn.running_mean"]) * scale


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(DEV).to(torch.bfloat16)


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_dwconv_matches_reference_conv_golden(ops, k):
    g = np.load(os.path.join(GOLDEN, "dwconv_ref.npz"))
    w, b = _fold(g, f"dw{k}_")
    c = w.shape[0]
    x = _nhwc(torch.from_numpy(g[f"dw{k}_x"]))
    y = torch.zeros(x.shape[0], x.shape[1], x.shape[2], c + 8, device=DEV, dtype=torch.bfloat16)     # channel slice of a wider buffer
    ops.dwconv(x, w.reshape(c, k * k).t().contiguous().to(DEV), b.to(DEV), y[..., :c], k)
    want = torch.from_numpy(g[f"dw{k}_y"]).permute(0, 2, 3, 1)
    assert rel_l2(y[..., :c], want) < 4e-3
    assert bool((y[..., c:] == 0).all())                       # the neighbouring slice is untouched


def test_dwconv_module_matches_reference_conv_golden():
    """Through the drop-in module API: Conv(c, c, k, 1, k//2, groups=c) with the reference's state_dict."""
    from yolo_ms_b200.modules import Conv
    g = np.load(os.path.join(GOLDEN, "dwconv_ref.npz"))
    for k in (3, 7):
        x = torch.from_numpy(g[f"dw{k}_x"])
        c = x.shape[1]
        m = Conv(c, c, k, 1, k // 2, groups=c)
        m.load_state_dict({n[len(f"dw{k}_"):]: torch.from_numpy(g[n]) for n in g.files
                           if n.startswith(f"dw{k}_") and n[len(f"dw{k}_"):].split(".")[0] in ("conv", "bn")}, strict=True)
        m = m.to(DEV).eval()
        assert rel_l2(m(x.to(DEV)), g[f"dw{k}_y"]) < 4e-3


def _layer_operands(g, tag):
    k = int(g[f"ms{tag}_k"])
    w1, b1 = _fold(g, f"ms{tag}_pw1.")
    wd, bd = _fold(g, f"ms{tag}_dw.")
    w2, b2 = _fold(g, f"ms{tag}_pw2.")
    e_ch, c = w1.shape[0], w1.shape[1]
    x = _nhwc(torch.from_numpy(g[f"ms{tag}_x"]))
    x2 = _nhwc(torch.from_numpy(g[f"ms{tag}_x2"])) if f"ms{tag}_x2" in g.files else None
    w1m = w1.reshape(e_ch, c)
    if x2 is not None:
        w1m = torch.cat([w1m, w1m], 1)                          # conv(x + x2): K-concatenation with repeated weights
    return dict(k=k, c=c, e_ch=e_ch, x=x, x2=x2, w1=w1m.contiguous().to(DEV).to(torch.bfloat16), b1=b1.to(DEV),
                wd=wd.reshape(e_ch, k * k).t().contiguous().to(DEV), bd=bd.to(DEV),
                w2=w2.reshape(c, e_ch).contiguous().to(DEV).to(torch.bfloat16), b2=b2.to(DEV))


def _run_layer(ops, o, mode):
    """mode 2: one kernel; mode 1: pw1 conv plan + (dw -> pw2) kernel; mode 0: three launches."""
    x, x2 = o["x"], o["x2"]
    b, h, w, c = x.shape
    y = torch.zeros(b, h, w, c + 8, device=DEV, dtype=torch.bfloat16)
    if mode == 2:
        ops.MsLayerPlan(2, y[..., :c], o["k"], o["wd"], o["bd"], x=x, x2=x2, w1=o["w1"], bias1=o["b1"], w2=o["w2"], bias2=o["b2"]).run()
    else:
        e = torch.empty(b, h, w, o["e_ch"], device=DEV, dtype=torch.bfloat16)
        ops.ConvPlan(x, o["w1"].unsqueeze(0).contiguous(), o["b1"], e, ksize=1, x2=x2).run()
        if mode == 1:
            ops.MsLayerPlan(1, y[..., :c], o["k"], o["wd"], o["bd"], e=e, w2=o["w2"], bias2=o["b2"]).run()
        else:
            d = torch.empty_like(e)
            ops.MsLayerPlan(0, d, o["k"], o["wd"], o["bd"], e=e).run()
            ops.ConvPlan(d, o["w2"].unsqueeze(0).contiguous(), o["b2"], y[..., :c], ksize=1).run()
    torch.cuda.synchronize()
    assert bool((y[..., c:] == 0).all())
    return y[..., :c]


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_ms_layer_modes_match_reference_golden(ops, tag):
    from yolo_ms_b200 import YmsError
    g = np.load(os.path.join(GOLDEN, "dwconv_ref.npz"))
    o = _layer_operands(g, tag)
    want = torch.from_numpy(g[f"ms{tag}_y"]).permute(0, 2, 3, 1)
    got = {}
    for mode in (0, 1, 2):
        try:
            got[mode] = _run_layer(ops, o, mode)
        except YmsError as err:                                # mode 2 only fits small layers (include/yms_b200.h)
            assert mode == 2 and "failed (code -2)" in str(err), err
            continue
        assert rel_l2(got[mode], want) < 1e-2, (tag, mode)
    assert 0 in got and 1 in got
    if tag in ("a", "b"):
        assert 2 in got, "the c <= 64 layers must take the fully fused mode"
    for mode in (1, 2):
        if mode in got:
            assert rel_l2(got[mode], got[0]) < 3e-3, (tag, mode)


@pytest.mark.parametrize("k,c,h,w,b,two", [(3, 32, 160, 160, 4, True), (3, 64, 80, 80, 8, False), (5, 64, 40, 40, 8, True),
                                           (5, 128, 40, 40, 16, True), (7, 256, 20, 20, 32, False), (7, 96, 24, 40, 2, False),
                                           (3, 48, 35, 50, 3, True)])
def test_ms_layer_persistent_tiles_vs_torch(ops, k, c, h, w, b, two):
    """Bench-sized maps (more tiles than SMs, several chunks per tile, maps that are not tile multiples) against a torch fp32
    evaluation of the same bf16-rounded operands with the kernel's storage contract (e and d rounded to bf16)."""
    from yolo_ms_b200 import YmsError
    g = torch.Generator().manual_seed(k * 1000 + c)
    bf = lambda t: t.to(torch.bfloat16).float()
    e_ch = 2 * c
    x = bf(torch.randn(b, c, h, w, generator=g))
    x2 = bf(torch.randn(b, c, h, w, generator=g)) if two else None
    w1 = bf(torch.randn(e_ch, c, generator=g) / c ** 0.5); b1 = torch.randn(e_ch, generator=g) * 0.2
    wd = torch.randn(e_ch, 1, k, k, generator=g) / k; bd = torch.randn(e_ch, generator=g) * 0.2
    w2 = bf(torch.randn(c, e_ch, generator=g) / e_ch ** 0.5); b2 = torch.randn(c, generator=g) * 0.2
    xd, x2d = x.to(DEV), (x2.to(DEV) if two else None)
    with torch.no_grad():
        s = xd + x2d if two else xd
        e = bf(F.silu(F.conv2d(s, w1.to(DEV).view(e_ch, c, 1, 1), b1.to(DEV))))
        d = bf(F.silu(F.conv2d(e, wd.to(DEV), bd.to(DEV), padding=k // 2, groups=e_ch)))
        want = F.silu(F.conv2d(d, w2.to(DEV).view(c, e_ch, 1, 1), b2.to(DEV))).permute(0, 2, 3, 1)
    o = dict(k=k, c=c, e_ch=e_ch, x=_nhwc(x), x2=_nhwc(x2) if two else None,
             w1=(torch.cat([w1, w1], 1) if two else w1).contiguous().to(DEV).to(torch.bfloat16), b1=b1.to(DEV),
             wd=wd.reshape(e_ch, k * k).t().contiguous().to(DEV), bd=bd.to(DEV),
             w2=w2.contiguous().to(DEV).to(torch.bfloat16), b2=b2.to(DEV))
    ran = 0
    for mode in (0, 1, 2):
        try:
            got = _run_layer(ops, o, mode)
        except YmsError:
            assert mode == 2
            continue
        ran += 1
        assert rel_l2(got, want) < 6e-3, (mode, rel_l2(got, want))
    assert ran >= 2


def test_ms_layer_is_deterministic(ops):
    """Same plan, same inputs -> same bits (static tile schedule, no atomics)."""
    g = np.load(os.path.join(GOLDEN, "dwconv_ref.npz"))
    o = _layer_operands(g, "a")
    a = _run_layer(ops, o, 2).clone()
    for _ in range(3):
        assert torch.equal(_run_layer(ops, o, 2), a)
