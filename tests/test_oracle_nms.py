"""Pin the NMS / post-process restatements against torchvision (the reference's third-party
dependency, present in this image) and against the committed golden vectors."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import postprocess as P

CASES = ("uniform", "clustered", "ties", "degenerate", "exact_thr")
THRS = (0.45, 0.5, 1.0 / 3.0)


@pytest.mark.parametrize("name", CASES)
def test_nms_restatements_match_golden(name):
    g = np.load(os.path.join(GOLDEN, "nms_cases.npz"))
    boxes, scores = g[f"{name}_boxes"], g[f"{name}_scores"]
    for thr in THRS:
        want = g[f"{name}_keep_{thr:.4f}"]
        assert np.array_equal(P.greedy_nms_numpy(boxes, scores, thr), want), (name, thr, "numpy")
        assert np.array_equal(P.greedy_nms_c(boxes, scores, thr), want), (name, thr, "C")


def test_nms_restatements_match_installed_torchvision():
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 33, 257, 1000):
        xy = rng.uniform(0, 100, (n, 2)); wh = rng.uniform(0, 40, (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = (np.round(rng.uniform(0, 1, n) * 32) / 32).astype(np.float32)
        want = P.torchvision_nms(boxes, scores, 0.5)
        assert np.array_equal(P.greedy_nms_numpy(boxes, scores, 0.5), want)
        assert np.array_equal(P.greedy_nms_c(boxes, scores, 0.5), want)


def test_postprocess_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "post_n.npz"))
    pred = g["pred"]
    for tag in ("a", "b"):
        conf, iou = g[f"thr_{tag}"]
        for i in range(pred.shape[0]):
            for fn in (P.greedy_nms_numpy, P.greedy_nms_c):
                keep, boxes, scores, labels = P.postprocess_image(pred[i], conf, iou, fn)
                assert np.array_equal(keep, g[f"keep_{tag}{i}"])
                assert np.array_equal(boxes, g[f"boxes_{tag}{i}"])
                assert np.array_equal(scores, g[f"scores_{tag}{i}"])
                assert np.array_equal(labels, g[f"labels_{tag}{i}"])
            # the fused C class-NMS agrees with the per-class loop
            b, s, l = P.select_candidates(pred[i])
            assert np.array_equal(P.class_nms_c(b, s, l, conf, iou), g[f"keep_{tag}{i}"])


def test_class_nms_edge_cases():
    z4 = np.zeros((0, 4), np.float32); z = np.zeros((0,), np.float32)
    assert P.class_nms(z4, z, np.zeros((0,), np.int64), 0.25, 0.45).size == 0
    assert P.class_nms_c(z4, z, np.zeros((0,), np.int32), 0.25, 0.45).size == 0
    # score exactly at conf is dropped (strict >), identical boxes of different classes both survive
    boxes = np.array([[0, 0, 10, 10]] * 3, np.float32)
    scores = np.array([0.25, 0.9, 0.8], np.float32)
    labels = np.array([0, 1, 2])
    assert P.class_nms(boxes, scores, labels, 0.25, 0.45).tolist() == [1, 2]
    assert P.class_nms_c(boxes, scores, labels, 0.25, 0.45).tolist() == [1, 2]
