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


def reference_postprocess(pred_single, conf_thresh, iou_thresh_nms):
    """tools/test.py:166-218 executed with torch + the real torchvision nms; returns the
    anchor indices the reference's final tensors correspond to."""
    from torchvision.ops import nms
    x_center, y_center, width, height = pred_single[:, :4].T
    x1 = x_center - width / 2
    y1 = y_center - height / 2
    x2 = x_center + width / 2
    y2 = y_center + height / 2
    boxes = torch.stack((x1, y1, x2, y2), dim=1)
    scores, class_indices = torch.max(pred_single[:, 4:], dim=1)
    conf_mask = scores > conf_thresh
    anchor_ids = torch.arange(pred_single.shape[0])[conf_mask]
    b, s, c = boxes[conf_mask], scores[conf_mask], class_indices[conf_mask]
    keep_ids, keep_boxes, keep_scores, keep_labels = [], [], [], []
    for cls_idx in torch.unique(c):
        cm = c == cls_idx
        cb, cs = b[cm], s[cm]
        if cb.shape[0] == 0:
            continue
        keep = nms(cb, cs, iou_thresh_nms)
        keep_ids.append(anchor_ids[cm][keep])
        keep_boxes.append(cb[keep])
        keep_scores.append(cs[keep])
        keep_labels.append(torch.full_like(cs[keep], fill_value=cls_idx.item(), dtype=torch.long))
    if not keep_ids:
        z = torch.zeros(0, dtype=torch.long)
        return z, torch.zeros(0, 4), torch.zeros(0), z
    return torch.cat(keep_ids), torch.cat(keep_boxes), torch.cat(keep_scores), torch.cat(keep_labels)


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
