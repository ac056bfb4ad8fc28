"""checkpoint ingestion (every layout the reference accepts) and the batched inference driver with the
reference's JSON record format."""
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from yolo_ms_b200.tools.utils import extract_state_dict, load_config


def test_extract_state_dict_accepts_every_reference_layout():
    sd = OrderedDict(a=torch.zeros(1), b=torch.ones(2))
    mod = OrderedDict(("module." + k, v) for k, v in sd.items())
    for ck in (sd, {"model": sd}, {"state_dict": sd}, mod, {"model": mod}, {"state_dict": mod, "epoch": 3}):
        out = extract_state_dict(ck)
        assert list(out.keys()) == ["a", "b"]
    with pytest.raises(TypeError):
        extract_state_dict([1, 2, 3])


def test_load_config_and_image_listing(tmp_path):
    from yolo_ms_b200.tools.test import list_images
    cfg = tmp_path / "c.yaml"
    cfg.write_text("model:\n  architecture: n\n  input_size: [64, 96]\ndataset:\n  num_classes: 3\n")
    c = load_config(str(cfg))
    assert c["model"]["architecture"] == "n" and c["model"]["input_size"] == [64, 96]
    for name in ("a.jpg", "b.png", "c.txt"):
        (tmp_path / name).write_bytes(b"x")
    assert sorted(os.path.basename(p) for p in list_images(str(tmp_path))) == ["a.jpg", "b.png"]
    with pytest.raises(FileNotFoundError):
        list_images(str(tmp_path / "missing"))


@pytest.mark.gpu
def test_checkpoint_round_trip_and_driver_json(tmp_path, monkeypatch):
    from PIL import Image
    from yolo_ms_b200 import engine
    monkeypatch.setattr(engine, "AUTOTUNE", False)      # two independently compiled programs must pick the same kernels
    from oracle import weights as W
    from yolo_ms_b200 import YOLOv8, ops
    from yolo_ms_b200.preprocess import preprocess_batch
    from yolo_ms_b200.tools.test import test as run_test
    from yolo_ms_b200.tools.utils import load_checkpoint

    sd = W.calibrated_state_dict("n", seed=1)
    ck = tmp_path / "model.pt"
    torch.save({"model": OrderedDict(("module." + k, v) for k, v in sd.items()), "epoch": 7}, ck)   # DataParallel + wrapper
    m = YOLOv8(version="n", num_classes=80)
    missing, unexpected = load_checkpoint(m, str(ck), strict=True)
    assert not missing and not unexpected
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k

    names = [f"thing{i}" for i in range(80)]
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(json.dumps({"device": "cuda", "model": {"architecture": "n", "input_size": [128, 160]},
                               "dataset": {"num_classes": 80, "class_names": names}}))       # JSON is YAML
    src = tmp_path / "imgs"
    src.mkdir()
    rng = np.random.default_rng(0)
    sizes = {"a": (97, 211), "b": (300, 180), "c": (128, 160)}
    raw = {}
    for k, (h, w) in sizes.items():
        raw[k] = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        Image.fromarray(raw[k]).save(src / f"{k}.png")
    out = tmp_path / "out"
    written = run_test(str(cfg), str(ck), str(src), str(out), 0.25, 0.45, batch_size=1)
    assert sorted(os.path.basename(p) for p in written) == ["a_detections.json", "b_detections.json", "c_detections.json"]

    # independent path: same uint8 batch through detect() + the reference's rescale / rounding arithmetic
    m = m.cuda().eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    for k, (h, w) in sizes.items():
        recs = json.load(open(out / f"{k}_detections.json"))
        batch = preprocess_batch([raw[k]], (128, 160))
        boxes, scores, labels, keep, count = m.detect(batch, 0.25, 0.45)
        n = int(count[0])
        idx = keep[0, :n].long()
        bx = boxes[0, idx].cpu().clone()
        bx[:, 0] *= w / 160; bx[:, 1] *= h / 128; bx[:, 2] *= w / 160; bx[:, 3] *= h / 128      # tools/test.py:222-229
        assert len(recs) == n and n > 0
        for r, b, s, l in zip(recs, bx.tolist(), scores[0, idx].cpu().tolist(), labels[0, idx].cpu().tolist()):
            assert r["box_xyxy"] == [round(c, 2) for c in b]
            assert r["score"] == round(s, 4) and r["class_id"] == l and r["class_name"] == names[l]
        assert set(recs[0].keys()) == {"box_xyxy", "score", "class_id", "class_name"}


# ----------------------------------------------------------------------------------------------
# the reference's own inference app as the pin (tests/golden/tools_test_sample.json = the records written by
# yolov8.tools.test.test() of the reference on its fixture yolov8/test/sample.png, oracle/make_golden.py::dump_tools_test)
# ----------------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tools_golden():
    from oracle import weights as W
    g = json.load(open(os.path.join(GOLD, "tools_test_sample.json")))
    rgb = np.load(os.path.join(GOLD, "tools_test_sample_image.npz"))["rgb"].copy()
    bn = np.load(os.path.join(GOLD, "tools_test_sample_bn.npz"))
    sd = W.make_state_dict(W.load_manifest("n"), seed=5)
    for k in bn.files:
        sd[k] = torch.from_numpy(bn[k].copy())
    return g, rgb, sd


def _matched(ref, got, iou_min=0.9, score_tol=0.05):
    m = 0
    for a in ref:
        for b in got:
            if a["class_id"] != b["class_id"]:
                continue
            A, B = a["box_xyxy"], b["box_xyxy"]
            ix = max(0.0, min(A[2], B[2]) - max(A[0], B[0])); iy = max(0.0, min(A[3], B[3]) - max(A[1], B[1]))
            inter = ix * iy
            u = (A[2] - A[0]) * (A[3] - A[1]) + (B[2] - B[0]) * (B[3] - B[1]) - inter
            if u > 0 and inter / u >= iou_min and abs(a["score"] - b["score"]) < score_tol:
                m += 1
                break
    return m


def test_oracle_reproduces_the_reference_apps_records_exactly():
    """Pre-processing (PIL bilinear resize, ToTensor, Normalize), fp32 forward, post-process, the rescale to the original size
    and the rounding of the JSON records -- restated with the oracle -- give the reference app's 160 records bit for bit.
    This pins every non-kernel step the GPU driver (yolo_ms_b200/tools/test.py) re-implements."""
    from PIL import Image
    from oracle import postprocess as PP
    from oracle import yolov8_oracle as O
    g, rgb, sd = _tools_golden()
    ih, iw = g["config"]["input_size"]
    oh, ow = rgb.shape[:2]
    im = np.asarray(Image.fromarray(rgb).resize((iw, ih), Image.BILINEAR)).copy()     # what T.Resize does to a PIL image
    x = torch.from_numpy(im).permute(2, 0, 1).float().div(255)
    x = ((x - torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)).unsqueeze(0)
    with torch.no_grad():
        pred = O.forward(sd, x)
    keep, b, s, l = PP.postprocess_image(pred[0].numpy(), g["config"]["conf_thresh"], g["config"]["iou_thresh_nms"], PP.greedy_nms_c)
    sx, sy = np.float32(ow / iw), np.float32(oh / ih)
    recs = [{"box_xyxy": [round(float(v), 2) for v in (np.float32(bb[0]) * sx, np.float32(bb[1]) * sy, np.float32(bb[2]) * sx, np.float32(bb[3]) * sy)],
             "score": round(float(ss), 4), "class_id": int(ll), "class_name": f"thing{int(ll)}"} for bb, ss, ll in zip(b, s, l)]
    assert recs == g["records"]


@pytest.mark.gpu
def test_gpu_driver_against_the_reference_apps_records():
    """The batched GPU driver on the same image and weights.  The model runs in bf16: the CPU oracle under the same numeric
    contract matches 149 of the reference's 160 records (same class, IoU >= 0.9, score within 0.05), so the gate is 1.5 x that
    miss rate; the record format, class names and ordering (class ascending, score descending) are exact."""
    from yolo_ms_b200 import YOLOv8
    from yolo_ms_b200.tools.test import detect_images
    g, rgb, sd = _tools_golden()
    m = YOLOv8(version="n", num_classes=80)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    names = [f"thing{i}" for i in range(80)]
    recs = detect_images(m, [rgb], tuple(g["config"]["input_size"]), g["config"]["conf_thresh"], g["config"]["iou_thresh_nms"], names)[0]
    ref = g["records"]
    assert abs(len(recs) - len(ref)) <= 16
    assert _matched(ref, recs) >= len(ref) - 17
    assert all(set(r.keys()) == {"box_xyxy", "score", "class_id", "class_name"} and r["class_name"] == names[r["class_id"]] for r in recs)
    cls = [r["class_id"] for r in recs]
    assert cls == sorted(cls)
    for c in set(cls):
        sc = [r["score"] for r in recs if r["class_id"] == c]
        assert sc == sorted(sc, reverse=True)
