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
