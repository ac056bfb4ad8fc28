"""Generate tests/golden/* from the REAL reference (/root/reference) and torchvision.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only runnable in the build container
(the GPU box has no /root/reference).  Run:  python -m oracle.make_golden

Outputs (all small, committed):
  manifest_{n,s,m}.json        (key, shape) list of the reference state_dict
  manifest_{n,s,m}_ms.json     derived manifest of the repo-local MS-Block variant
  model_{n,s}.npz              reference forward on seeded weights/images: pred, raw head
                               tensors, backbone/neck features
  post_n.npz                   reference post-process (tools/test.py:166-218 with the real
                               torchvision.ops.nms) on a reference prediction
  nms_cases.npz                torchvision.ops.nms keep lists on adversarial box sets
  tools_test_sample.json       the detection records written by the reference's inference app test() (tools/test.py:63-276)
  dwconv_ref.npz               the reference's own Conv unit run as a depthwise layer, Conv(c, c, k, 1, k//2, groups=c)
                               (components.py:69-77; the reference never instantiates it that way, but the class supports it),
                               k = 3/5/7/9, and the MS-Block branch layer pw1 -> dw -> pw2 composed of three reference Conv units
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import weights as W                      # noqa: E402
from oracle import yolov8_oracle as O                # noqa: E402

GOLD = W.GOLDEN_DIR


def ref_model(version, sd, nc=80):
    from yolov8.yolov8 import YOLOv8                  # the real reference
    m = YOLOv8(version=version, num_classes=nc)
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])   # tools/test.py:110-111
    return m


def dump_manifests():
    from yolov8.yolov8 import YOLOv8
    for v in ("n", "s", "m"):
        m = YOLOv8(version=v, num_classes=80)
        man = [(k, list(t.shape)) for k, t in m.state_dict().items()]
        with open(os.path.join(GOLD, f"manifest_{v}.json"), "w") as f:
            json.dump(man, f)
        ms = O.ms_manifest_from_c2f(dict(man))
        with open(os.path.join(GOLD, f"manifest_{v}_ms.json"), "w") as f:
            json.dump([(k, s) for k, s in ms.items()], f)
        print(v, len(man), "entries;", sum(int(np.prod(s)) for _, s in man if len(s)), "elements")


@torch.no_grad()
def dump_model(version, batch, h, w, seed):
    sd = W.calibrated_state_dict(version, seed=seed)
    m = ref_model(version, sd)
    x = W.make_images(batch, h, w, seed=7)
    feats = m.backbone(x)
    nk = m.neck(*feats)
    m.head.training = True                            # test_model.py:235 toggles it this way
    raw = m.head([t.clone() for t in nk])
    m.head.training = False
    pred = m(x)
    out = {"pred": pred.numpy(), "x_shape": np.array(x.shape)}
    for i in range(3):
        out[f"p{i}"] = feats[i].numpy()
        out[f"n{i}"] = nk[i].numpy()
        out[f"raw{i}"] = raw[i].numpy()
    np.savez_compressed(os.path.join(GOLD, f"model_{version}.npz"), **out)
    print("model", version, {k: v.shape for k, v in out.items()},
          "feat std", [float(t.std()) for t in feats], "raw std", [float(t.std()) for t in raw])
    return pred


from oracle.postprocess import torch_postprocess_image as reference_postprocess   # noqa: E402  (tools/test.py:166-218 with the real torchvision nms)


@torch.no_grad()
def dump_post():
    sd = W.calibrated_state_dict("n", seed=3)
    m = ref_model("n", sd)
    x = W.make_images(2, 160, 192, seed=11)
    pred = m(x)
    out = {"pred": pred.numpy()}
    for conf, iou, tag in ((0.25, 0.45, "a"), (0.6, 0.5, "b")):
        for i in range(pred.shape[0]):
            ids, bx, sc, lb = reference_postprocess(pred[i], conf, iou)
            out[f"keep_{tag}{i}"] = ids.numpy()
            out[f"boxes_{tag}{i}"] = bx.numpy()
            out[f"scores_{tag}{i}"] = sc.numpy()
            out[f"labels_{tag}{i}"] = lb.numpy()
            print("post", tag, i, "kept", ids.numel(), "classes", int(torch.unique(lb).numel()))
        out[f"thr_{tag}"] = np.array([conf, iou])
    np.savez_compressed(os.path.join(GOLD, "post_n.npz"), **out)


def nms_case(name, seed):
    g = np.random.default_rng(seed)
    if name == "uniform":
        n = 1500
        xy = g.uniform(0, 600, (n, 2)); wh = g.uniform(4, 64, (n, 2))
        sc = g.uniform(0, 1, n)
    elif name == "clustered":
        c = g.uniform(50, 550, (30, 2)); n = 1500
        xy = np.repeat(c, 50, 0) + g.normal(0, 4, (n, 2)); wh = 40 + g.normal(0, 3, (n, 2))
        sc = g.uniform(0, 1, n)
    elif name == "ties":
        n = 1200
        xy = g.uniform(0, 200, (n, 2)); wh = g.uniform(10, 60, (n, 2))
        sc = np.round(g.uniform(0, 1, n) * 16) / 16          # heavy score ties
    elif name == "degenerate":
        n = 600
        xy = np.round(g.uniform(0, 64, (n, 2))); wh = np.round(g.uniform(0, 8, (n, 2)))   # zero-area boxes
        sc = np.round(g.uniform(0, 1, n) * 8) / 8
    elif name == "exact_thr":
        # integer boxes on a lattice: many pairs have IoU exactly 1/3, 1/2, ... (thr hit exactly)
        n = 800
        xy = g.integers(0, 12, (n, 2)).astype(np.float64) * 2; wh = g.integers(1, 4, (n, 2)).astype(np.float64) * 2
        sc = g.uniform(0, 1, n)
    else:
        raise KeyError(name)
    boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    return boxes, sc.astype(np.float32)


def dump_nms():
    from torchvision.ops import nms
    out = {}
    for ci, name in enumerate(("uniform", "clustered", "ties", "degenerate", "exact_thr")):
        boxes, sc = nms_case(name, 100 + ci)
        out[f"{name}_boxes"] = boxes
        out[f"{name}_scores"] = sc
        for thr in (0.45, 0.5, 1.0 / 3.0):
            keep = nms(torch.from_numpy(boxes), torch.from_numpy(sc), thr).numpy()
            out[f"{name}_keep_{thr:.4f}"] = keep
            print("nms", name, thr, "kept", keep.size, "of", boxes.shape[0])
    np.savez_compressed(os.path.join(GOLD, "nms_cases.npz"), **out)


def _randomize_bn(mod, g):
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            m.weight.data = torch.rand(n, generator=g) * 0.8 + 0.6
            m.bias.data = torch.randn(n, generator=g) * 0.2
            m.running_mean.data = torch.randn(n, generator=g) * 0.2
            m.running_var.data = torch.rand(n, generator=g) * 0.8 + 0.6


@torch.no_grad()
def dump_dwconv():
    """Depthwise unit and MS-Block branch layer computed by the REFERENCE's Conv class (fp32, eval)."""
    from yolov8.model.components import Conv            # the real reference
    out = {}
    g = torch.Generator().manual_seed(123)
    bf = lambda t: t.to(torch.bfloat16).float()          # inputs / conv weights are bf16-representable (the GPU contract stores them so)
    for k, (c, h, w) in ((3, (64, 20, 24)), (5, (24, 19, 37)), (7, (72, 20, 20)), (9, (16, 12, 33))):
        m = Conv(c, c, k, 1, k // 2, groups=c).eval()
        m.conv.weight.data = torch.randn(c, 1, k, k, generator=g) / k
        _randomize_bn(m, g)
        x = bf(torch.randn(1, c, h, w, generator=g))
        out[f"dw{k}_x"] = x.numpy()
        out[f"dw{k}_y"] = m(x.clone()).numpy()
        for name, t in m.state_dict().items():
            out[f"dw{k}_{name}"] = t.numpy()
    # branch layer: x (+ x2) -> Conv(c, 2c, 1) -> Conv(2c, 2c, k, groups=2c) -> Conv(2c, c, 1)
    for tag, (k, c, h, w, two) in (("a", (3, 32, 24, 40, True)), ("b", (3, 64, 17, 21, False)), ("c", (5, 128, 12, 20, True)),
                                   ("d", (7, 48, 9, 13, False))):
        pw1, dw, pw2 = Conv(c, 2 * c, 1, 1, 0).eval(), Conv(2 * c, 2 * c, k, 1, k // 2, groups=2 * c).eval(), Conv(2 * c, c, 1, 1, 0).eval()
        pw1.conv.weight.data = bf(torch.randn(2 * c, c, 1, 1, generator=g) / c ** 0.5)
        dw.conv.weight.data = torch.randn(2 * c, 1, k, k, generator=g) / k
        pw2.conv.weight.data = bf(torch.randn(c, 2 * c, 1, 1, generator=g) / (2 * c) ** 0.5)
        for m in (pw1, dw, pw2):
            _randomize_bn(m, g)
        x = bf(torch.randn(1, c, h, w, generator=g))
        x2 = bf(torch.randn(1, c, h, w, generator=g)) if two else None
        y = pw2(dw(pw1(x + x2 if two else x.clone())))
        out[f"ms{tag}_x"] = x.numpy()
        if two:
            out[f"ms{tag}_x2"] = x2.numpy()
        out[f"ms{tag}_y"] = y.numpy()
        out[f"ms{tag}_k"] = np.array(k)
        for nm, m in (("pw1", pw1), ("dw", dw), ("pw2", pw2)):
            for name, t in m.state_dict().items():
                out[f"ms{tag}_{nm}.{name}"] = t.numpy()
    np.savez_compressed(os.path.join(GOLD, "dwconv_ref.npz"), **out)
    print("dwconv_ref", {k: v.shape for k, v in out.items() if k.endswith("_y")})


def dump_tools_test():
    """The REAL inference app of the reference, yolov8.tools.test.test() (tools/test.py:63-276), run on the reference's own
    fixture image yolov8/test/sample.png (137x138 RGBA) with a seeded checkpoint: commits the JSON records it writes
    (tests/golden/tools_test_sample.json) plus the decoded RGB pixels of the fixture (the GPU box has no /root/reference)."""
    import tempfile
    import yaml
    from PIL import Image
    from yolov8.tools.test import test as ref_test          # the real reference app
    src = os.path.join(REF, "yolov8", "test", "sample.png")
    with tempfile.TemporaryDirectory() as td:
        ck = os.path.join(td, "best.pt")
        # seeded weights whose BN statistics are calibrated ON this image (+ one synthetic one): logits stay O(1) and scores
        # spread over (0, 1) instead of saturating at 1.0
        import torchvision.transforms as T
        tf = T.Compose([T.Resize((160, 192)), T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        xs = tf(Image.open(src).convert("RGB")).unsqueeze(0)
        sd = W.make_state_dict(W.load_manifest("n"), seed=5)
        with torch.no_grad():
            O.calibrate_bn(sd, torch.cat([xs, W.make_images(1, 160, 192, seed=55)]))
        torch.save(sd, ck)
        np.savez_compressed(os.path.join(GOLD, "tools_test_sample_bn.npz"),
                            **{k: v.numpy() for k, v in sd.items() if k.endswith(("running_mean", "running_var"))})
        cfg = os.path.join(td, "cfg.yaml")
        names = [f"thing{i}" for i in range(80)]
        with open(cfg, "w") as f:
            yaml.safe_dump({"device": "cpu", "model": {"architecture": "n", "input_size": [160, 192]},
                            "dataset": {"num_classes": 80, "class_names": names}}, f)
        out = os.path.join(td, "out")
        ref_test(cfg, ck, src, out, conf_thresh=0.25, iou_thresh_nms=0.45)
        with open(os.path.join(out, "sample_detections.json")) as f:
            recs = json.load(f)
    with open(os.path.join(GOLD, "tools_test_sample.json"), "w") as f:
        json.dump({"config": {"architecture": "n", "input_size": [160, 192], "num_classes": 80, "weights": "make_state_dict(manifest n, seed 5) + BN running stats of tools_test_sample_bn.npz",
                              "conf_thresh": 0.25, "iou_thresh_nms": 0.45},
                   "produced_by": "yolov8.tools.test.test() of the reference on yolov8/test/sample.png (oracle/make_golden.py::dump_tools_test)",
                   "records": recs}, f)
    np.savez_compressed(os.path.join(GOLD, "tools_test_sample_image.npz"), rgb=np.asarray(Image.open(src).convert("RGB")))
    print("tools_test", len(recs), "records; first", recs[0] if recs else None)


def dump_bn_fixtures():
    """BN running statistics of the calibrated synthetic weights (seed 1) for bench.py / smoke():
    yolo_ms_b200/synth_bn/bn_{version}_{block}_seed1.npz."""
    from yolo_ms_b200 import synth
    os.makedirs(synth.BN_DIR, exist_ok=True)
    for v in ("n", "s", "m"):
        for blk in ("c2f", "ms"):
            sd = W.calibrated_state_dict(v, seed=1, block=blk)
            stats = {k: t.numpy() for k, t in sd.items() if k.endswith(("running_mean", "running_var"))}
            np.savez_compressed(synth.bn_fixture_path(v, blk, 1), **stats)
            print("bn fixture", v, blk, len(stats))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    dump_manifests()
    dump_bn_fixtures()
    dump_model("n", 2, 64, 96, seed=1)
    dump_model("s", 1, 64, 64, seed=2)
    dump_post()
    dump_nms()
    dump_dwconv()
    dump_tools_test()
