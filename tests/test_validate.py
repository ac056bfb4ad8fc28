"""mAP@0.5 bookkeeping (COCO protocol restated; parity unpinned -- torchmetrics is not available offline).
Known-answer cases computed by hand from the protocol."""
import numpy as np
import pytest
import torch

from yolo_ms_b200.tools.validate import MeanAveragePrecision50, box_iou_np


def _box(x, y, s=10.0):
    return [x, y, x + s, y + s]


def test_iou_matrix():
    a = np.array([[0, 0, 10, 10]], np.float32); b = np.array([[0, 0, 10, 10], [5, 0, 15, 10], [20, 20, 30, 30]], np.float32)
    assert np.allclose(box_iou_np(a, b), [[1.0, 1 / 3, 0.0]])


def test_perfect_detections_give_one():
    m = MeanAveragePrecision50()
    gt = {"boxes": np.array([_box(0, 0), _box(50, 50)], np.float32), "labels": np.array([0, 3])}
    m.update([{"boxes": gt["boxes"], "scores": np.array([0.9, 0.8], np.float32), "labels": gt["labels"]}], [gt])
    assert float(m.compute()["map_50"]) == pytest.approx(1.0)


def test_hand_computed_ap():
    # one class, 2 GT; detections by score: TP, FP, TP -> precision envelope 1.0 up to recall 0.5, 2/3 up to recall 1.0
    m = MeanAveragePrecision50()
    gt = {"boxes": np.array([_box(0, 0), _box(50, 50)], np.float32), "labels": np.array([1, 1])}
    pr = {"boxes": np.array([_box(0, 0), _box(200, 200), _box(51, 50)], np.float32),
          "scores": np.array([0.9, 0.8, 0.7], np.float32), "labels": np.array([1, 1, 1])}
    m.update([pr], [gt])
    assert float(m.compute()["map_50"]) == pytest.approx((51 * 1.0 + 50 * (2 / 3)) / 101, abs=1e-6)


def test_duplicates_are_false_positives_and_missing_classes_count():
    m = MeanAveragePrecision50()
    gt = {"boxes": np.array([_box(0, 0), _box(100, 100)], np.float32), "labels": np.array([0, 2])}
    pr = {"boxes": np.array([_box(0, 0), _box(1, 0)], np.float32), "scores": np.array([0.9, 0.6], np.float32), "labels": np.array([0, 0])}
    m.update([pr], [gt])                       # class 0: TP then duplicate FP -> AP 1.0; class 2: no detection -> AP 0
    r = m.compute()
    assert r["ap_per_class"][0] == pytest.approx(1.0) and r["ap_per_class"][2] == 0.0
    assert float(r["map_50"]) == pytest.approx(0.5)


def test_accumulates_over_images_and_accepts_tensors():
    m = MeanAveragePrecision50()
    for k in range(3):
        gt = {"boxes": torch.tensor([_box(10 * k, 0)]), "labels": torch.tensor([5])}
        pr = {"boxes": torch.tensor([_box(10 * k, 0)] if k < 2 else [_box(300, 300)]), "scores": torch.tensor([0.9 - 0.1 * k]), "labels": torch.tensor([5])}
        m.update([pr], [gt])
    # 3 GT; sorted: TP(.9) TP(.8) FP(.7): recall reaches 2/3 at precision 1 -> 67 of 101 thresholds (0..0.66) sampled at 1.0
    assert float(m.compute()["map_50"]) == pytest.approx(67 / 101, abs=1e-6)


@pytest.mark.gpu
def test_validate_epoch_runs_on_gpu_detections():
    from oracle import weights as W
    from yolo_ms_b200 import YOLOv8
    from yolo_ms_b200.tools.validate import validate_epoch
    m = YOLOv8(version="n", num_classes=80)
    m.load_state_dict(W.calibrated_state_dict("n", seed=1))
    m = m.cuda().eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    x = W.make_images(2, 128, 160, seed=3)
    # ground truth = the model's own detections of image 0 -> that image scores AP 1 for its classes
    boxes, scores, labels, keep, count = m.detect(x.cuda(), 0.25, 0.45)
    idx = keep[0, :int(count[0])].long()
    b = boxes[0, idx].cpu().numpy(); l = labels[0, idx].cpu().numpy()
    t = np.stack([np.zeros(len(l)), l, (b[:, 0] + b[:, 2]) / 2 / 160, (b[:, 1] + b[:, 3]) / 2 / 128, (b[:, 2] - b[:, 0]) / 160, (b[:, 3] - b[:, 1]) / 128], 1)
    cfg = {"model": {"input_size": [128, 160]}, "evaluation": {}}
    v = validate_epoch(m, [(x, torch.tensor(t, dtype=torch.float32))], torch.device("cuda"), cfg)
    assert 0.3 < v <= 1.0
