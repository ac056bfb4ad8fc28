"""The reference's inline post-process (yolov8/tools/test.py:166-218, duplicated at
yolov8/tools/train.py:63-113) as a callable: candidate selection + confidence filter +
class-aware NMS, batched, on the GPU (kernels yms_select_candidates + yms_nms_batched).

Output convention = the reference's final tensors per image: boxes xyxy (model input scale),
scores, int64 class indices, ordered class-ascending then score-descending.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops


@torch.no_grad()
def postprocess_batched(pred: torch.Tensor, conf_thresh: float = 0.25, iou_thresh: float = 0.45):
    """pred [B, A, 4+nc] fp32 (YOLOv8.forward eval output).  Returns the padded form:
    (boxes_xyxy [B,A,4], scores [B,A], labels int32 [B,A], keep int32 [B,A] (-1 padded), count [B])."""
    boxes, scores, labels = ops.select_candidates(pred)
    keep, count = ops.nms_batched(boxes, scores, labels, conf_thresh, iou_thresh, pred.shape[-1] - 4)
    return boxes, scores, labels, keep, count


@torch.no_grad()
def postprocess(pred: torch.Tensor, conf_thresh: float = 0.25, iou_thresh: float = 0.45
                ) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """Per image (final_boxes [K,4], final_scores [K], final_class_indices [K] int64), exactly the
    three tensors tools/test.py:216-218 builds.  Accepts [A,4+nc] (one image) or [B,A,4+nc]."""
    single = pred.dim() == 2
    if single:
        pred = pred.unsqueeze(0)
    boxes, scores, labels, keep, count = postprocess_batched(pred, conf_thresh, iou_thresh)
    counts = count.tolist()                               # the one device->host sync of the post-process
    out = []
    for i, k in enumerate(counts):
        idx = keep[i, :k].long()
        out.append((boxes[i, idx], scores[i, idx], labels[i, idx].long()))
    return out[0] if single else out
