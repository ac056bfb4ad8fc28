"""Pin the oracle's model restatement: against the committed goldens (produced by the REAL
reference, oracle/make_golden.py) everywhere, and against the live reference when
/root/reference is present (build container only)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, have_reference
from oracle import weights as W
from oracle import yolov8_oracle as O

# fp32 conv summation order differs between CPUs / mkldnn kernels: compare with a tolerance
# relative to the tensor's scale.
def close(a, b, tol=2e-4):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12)) < tol


@pytest.mark.parametrize("version,batch,h,w,seed", [("n", 2, 64, 96, 1), ("s", 1, 64, 64, 2)])
def test_oracle_matches_reference_golden(version, batch, h, w, seed):
    g = np.load(os.path.join(GOLDEN, f"model_{version}.npz"))
    sd = W.calibrated_state_dict(version, seed=seed)
    x = W.make_images(batch, h, w, seed=7)
    with torch.no_grad():
        r = O.forward(sd, x, return_parts=True)
    for i in range(3):
        assert close(r["p"][i], g[f"p{i}"]), f"backbone P{i + 3}"
        assert close(r["n"][i], g[f"n{i}"]), f"neck N{i + 3}"
        assert close(r["raw"][i], g[f"raw{i}"]), f"raw head {i}"
    pred = r["pred"]
    assert pred.shape == tuple(g["pred"].shape)
    assert close(pred[..., :4], g["pred"][..., :4], 5e-4)
    assert float((pred[..., 4:] - torch.from_numpy(g["pred"][..., 4:])).abs().max()) < 1e-3


def test_manifest_and_anchor_count():
    man = W.load_manifest("n")
    assert len(man) == 355                       # SURVEY 3.4 [measured]
    assert not any("stride" in k for k in man)   # head.stride is not a buffer
    sd = W.make_state_dict(man, seed=0)
    assert torch.equal(sd["head.dfl.conv.weight"].flatten(), torch.arange(16.0))
    with torch.no_grad():
        pred = O.forward(sd, torch.zeros(1, 3, 64, 64))
    assert pred.shape == (1, 64 + 16 + 4, 84)


def test_zero_stride_gives_zero_boxes():
    sd = W.make_state_dict(W.load_manifest("n"), seed=0)
    with torch.no_grad():
        pred = O.forward(sd, W.make_images(1, 64, 64), strides=(0.0, 0.0, 0.0))
    assert float(pred[..., :4].abs().max()) == 0.0       # yolov8_head.py:79 default stride


def test_decode_anchor_layout():
    # all-zero box logits -> uniform DFL -> dist 7.5 each side: cx,cy = anchor*stride, wh = 15*stride
    raw = [torch.zeros(1, 144, 4, 6), torch.zeros(1, 144, 2, 3), torch.zeros(1, 144, 1, 2)]
    pred = O.decode(raw, (8.0, 16.0, 32.0))
    assert pred.shape == (1, 24 + 6 + 2, 84)
    assert torch.allclose(pred[0, 0, :4], torch.tensor([4.0, 4.0, 120.0, 120.0]))
    assert torch.allclose(pred[0, 1, :2], torch.tensor([12.0, 4.0]))     # x fastest
    assert torch.allclose(pred[0, 6, :2], torch.tensor([4.0, 12.0]))
    assert torch.allclose(pred[0, 24, :4], torch.tensor([8.0, 8.0, 240.0, 240.0]))
    assert torch.allclose(pred[0, :, 4:], torch.full((32, 80), 0.5))


def test_ms_manifest_shapes_and_forward():
    man = W.load_manifest("n", "ms")
    assert "backbone.c2f_2.in_conv.conv.weight" in man
    assert man["backbone.c2f_8.branches.0.0.dw.conv.weight"][1:] == [1, 7, 7]
    sd = W.make_state_dict(man, seed=0)
    with torch.no_grad():
        pred = O.forward(sd, W.make_images(1, 64, 64))
    assert pred.shape == (1, 84, 84) and torch.isfinite(pred).all()


@pytest.mark.skipif(not have_reference(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("version", ["n", "s", "m"])
def test_oracle_matches_live_reference(version):
    sys.path.insert(0, REFERENCE)
    from yolov8.yolov8 import YOLOv8
    sd = W.calibrated_state_dict(version, seed=5)
    m = YOLOv8(version=version, num_classes=80)
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    x = W.make_images(2, 96, 128, seed=9)
    with torch.no_grad():
        ref = m(x)
        got = O.forward(sd, x)
        m.head.training = True
        raw_ref = m.head(list(m.neck(*m.backbone(x))))
        raw = O.forward(sd, x, return_parts=True)["raw"]
    assert got.shape == ref.shape
    assert close(got[..., :4], ref[..., :4], 1e-5)
    assert float((got[..., 4:] - ref[..., 4:]).abs().max()) < 1e-5
    for a, b in zip(raw, raw_ref):
        assert close(a, b, 1e-5)


def test_oracle_conv_unit_matches_reference_conv_as_depthwise_and_ms_layer():
    """tests/golden/dwconv_ref.npz was produced by the reference's own Conv class (components.py:69-77) run with groups=c and
    as the pw1 -> dw -> pw2 chain of an MS-Block branch layer (oracle/make_golden.py::dump_dwconv): pins the oracle's conv unit
    for the depthwise case, which the reference model itself never instantiates."""
    g = np.load(os.path.join(GOLDEN, "dwconv_ref.npz"))
    for k in (3, 5, 7, 9):
        sd = {"u." + n[len(f"dw{k}_"):]: torch.from_numpy(g[n]) for n in g.files if n.startswith(f"dw{k}_") and n[len(f"dw{k}_")] in "cb"}
        y = O.conv_unit(sd, "u", torch.from_numpy(g[f"dw{k}_x"]))
        assert float((y - torch.from_numpy(g[f"dw{k}_y"])).abs().max()) < 2e-5
    for tag in "abcd":
        sd = {n[len(f"ms{tag}_"):]: torch.from_numpy(g[n]) for n in g.files if n.startswith(f"ms{tag}_") and "." in n}
        x = torch.from_numpy(g[f"ms{tag}_x"])
        if f"ms{tag}_x2" in g.files:
            x = x + torch.from_numpy(g[f"ms{tag}_x2"])
        y = O.conv_unit(sd, "pw2", O.conv_unit(sd, "dw", O.conv_unit(sd, "pw1", x)))
        assert float((y - torch.from_numpy(g[f"ms{tag}_y"])).abs().max()) < 2e-5
