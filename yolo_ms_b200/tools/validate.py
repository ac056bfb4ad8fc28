"""Validation loop / mAP@0.5 on GPU detections, mirroring ``validate_epoch`` of the reference
(yolov8/tools/train.py:19-165): model -> class-aware NMS -> ``MeanAveragePrecision(iou_thresholds=[0.5])``.

The reference delegates the metric to torchmetrics (third party, not vendored, absent from this image together with its
pycocotools backend), so the COCO evaluation protocol is restated here -- **parity unpinned**: there is no reference
output to compare against offline.  Protocol (pycocotools ``COCOeval``, bbox, IoU 0.5, area "all", maxDets 100):
per (image, class) detections sorted by score (stable), the top 100 kept; each detection is matched to the unmatched
ground-truth box of highest IoU >= 0.5; per class, precision is made monotonically non-increasing and sampled at the
101 recall thresholds 0:0.01:1; AP = mean of the samples; mAP = mean over the classes that have ground truth.

The forward / decode / NMS run on the GPU (YOLOv8.detect); the metric bookkeeping is host-side numpy.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from ..ops import gather_detections

REC_THRS = np.linspace(0.0, 1.0, 101)
MAX_DETS = 100


def box_iou_np(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """IoU matrix of xyxy boxes [n,4] x [m,4] (float64, like pycocotools' maskUtils.iou on bboxes)."""
    a = a.astype(np.float64); b = b.astype(np.float64)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return np.zeros((a.shape[0], b.shape[0]))
    iw = np.clip(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0, None)
    ih = np.clip(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]), 0, None)
    inter = iw * ih
    ua = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ub = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    union = ua[:, None] + ub[None, :] - inter
    return np.where(union > 0, inter / np.where(union > 0, union, 1.0), 0.0)


class MeanAveragePrecision50:
    """Drop-in for ``MeanAveragePrecision(box_format='xyxy', iou_type='bbox', iou_thresholds=[0.5])`` as the reference
    uses it (train.py:40-46,142,148-150): ``update(preds, targets)`` with lists of dicts, ``compute()['map_50']``."""

    def __init__(self, iou_threshold: float = 0.5):
        self.thr = float(iou_threshold)
        self.records: Dict[int, List] = {}      # class -> list of (scores [k], matched [k]) per image
        self.npos: Dict[int, int] = {}

    @staticmethod
    def _np(t):
        return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)

    def update(self, preds: Sequence[dict], targets: Sequence[dict]) -> None:
        for p, t in zip(preds, targets):
            pb, ps, pl = self._np(p["boxes"]).reshape(-1, 4), self._np(p["scores"]).reshape(-1), self._np(p["labels"]).reshape(-1).astype(np.int64)
            tb, tl = self._np(t["boxes"]).reshape(-1, 4), self._np(t["labels"]).reshape(-1).astype(np.int64)
            for c in np.union1d(pl, tl):
                c = int(c)
                gt = tb[tl == c]
                self.npos[c] = self.npos.get(c, 0) + gt.shape[0]
                sel = np.nonzero(pl == c)[0]
                if sel.size == 0:
                    continue
                order = np.argsort(-ps[sel], kind="mergesort")[:MAX_DETS]
                dt, sc = pb[sel][order], ps[sel][order]
                ious = box_iou_np(dt, gt)
                taken = np.zeros(gt.shape[0], bool)
                matched = np.zeros(dt.shape[0], bool)
                for d in range(dt.shape[0]):
                    best, m = min(self.thr, 1 - 1e-10), -1
                    for g in range(gt.shape[0]):
                        if taken[g] or ious[d, g] < best:
                            continue
                        best, m = ious[d, g], g
                    if m >= 0:
                        taken[m] = True
                        matched[d] = True
                self.records.setdefault(c, []).append((sc, matched))

    def compute(self) -> dict:
        aps = {}
        for c, npos in self.npos.items():
            if npos == 0:
                continue
            recs = self.records.get(c, [])
            if recs:
                sc = np.concatenate([r[0] for r in recs]); mt = np.concatenate([r[1] for r in recs])
                order = np.argsort(-sc, kind="mergesort")
                tp = np.cumsum(mt[order]).astype(np.float64); fp = np.cumsum(~mt[order]).astype(np.float64)
                rc = tp / npos
                pr = tp / (tp + fp + np.spacing(1))
                for i in range(len(pr) - 1, 0, -1):
                    if pr[i] > pr[i - 1]:
                        pr[i - 1] = pr[i]
                inds = np.searchsorted(rc, REC_THRS, side="left")
                q = np.zeros(len(REC_THRS))
                ok = inds < len(pr)
                q[ok] = pr[inds[ok]]
                aps[c] = float(q.mean())
            else:
                aps[c] = 0.0
        m = float(np.mean(list(aps.values()))) if aps else -1.0
        return {"map_50": torch.tensor(m), "map": torch.tensor(m), "ap_per_class": aps}


@torch.no_grad()
def validate_epoch(model, val_loader, device, cfg, epoch_num=-1) -> float:
    """Same contract as the reference's ``validate_epoch`` (train.py:19-165): ``val_loader`` yields (images, targets) with
    images [B,3,H,W] normalised fp32 (or uint8 [B,H,W,3]) and targets [N,6] = (image index, class, cx, cy, w, h normalised);
    returns mAP@0.5.  The per-image Python post-process of the reference is one batched detect() here."""
    model.eval()
    ev = cfg.get("evaluation", {})
    conf, iou = ev.get("confidence_threshold", 0.25), ev.get("iou_threshold", 0.45)
    in_h, in_w = cfg["model"].get("input_size", [640, 640])
    metric = MeanAveragePrecision50()
    total, n_img = 0, 0
    for images, targets in val_loader:
        images = images.to(device)
        b = images.shape[0]
        n_img += b
        boxes, scores, labels, keep, count = model.detect(images, conf, iou)
        dets = gather_detections(boxes, scores, labels, keep, count, boxes.shape[1]).cpu().numpy()
        cnt = count.cpu().numpy()
        tg = targets.detach().cpu().numpy() if torch.is_tensor(targets) else np.asarray(targets)
        preds, gts = [], []
        for i in range(b):
            d = dets[i, :cnt[i]]
            total += int(cnt[i])
            preds.append({"boxes": d[:, :4], "scores": d[:, 4], "labels": d[:, 5].astype(np.int64)})
            g = tg[tg[:, 0] == i][:, 1:] if tg.size else np.zeros((0, 5), np.float32)
            cx, cy, w, h = g[:, 1] * in_w, g[:, 2] * in_h, g[:, 3] * in_w, g[:, 4] * in_h      # train.py:123-132
            gts.append({"boxes": np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1) if g.shape[0] else np.zeros((0, 4)),
                        "labels": g[:, 0].astype(np.int64)})
        metric.update(preds, gts)
    res = metric.compute()
    m = float(res["map_50"])
    print(f"--- Validation Summary ---\nProcessed {n_img} images.\nTotal Detections (after NMS & conf_thresh): {total}\n"
          f"Average Detections per Image: {total / max(n_img, 1):.2f}\nmAP@0.5: {m:.4f}")
    return m
