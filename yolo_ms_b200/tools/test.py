"""Batched inference driver: the reference's ``yolov8/tools/test.py`` (``test()``, :62-275)
with the per-image Python loop replaced by batches that stay on the GPU: decode on the host (PIL) -> uint8 upload ->
GPU resize (Pillow-exact) -> fused ToTensor/Normalize + forward + DFL decode + class-aware NMS -> boxes rescaled to the
original image size -> one ``<name>_detections.json`` per image with the reference's record format
(``box_xyxy`` rounded to 2, ``score`` to 4, ``class_id``, ``class_name``; tools/test.py:254-273).

Drawing the boxes on the images (tools/test.py:18-60,232-251) is host-side visualisation and out of scope here.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from ..modules import YOLOv8
from ..ops import gather_detections
from ..preprocess import preprocess_batch
from .utils import load_checkpoint, load_config

EXTENSIONS = ["*.jpg", "*.jpeg", "*.png", "*.bmp", "*.tif", "*.tiff"]     # tools/test.py:125


def list_images(source_path: str) -> List[str]:
    """tools/test.py:121-136."""
    if os.path.isdir(source_path):
        paths: List[str] = []
        for ext in EXTENSIONS:
            paths.extend(glob.glob(os.path.join(source_path, ext)))
        return paths
    if os.path.isfile(source_path):
        return [source_path]
    raise FileNotFoundError(f"Source path not found or is not a file/directory: {source_path}")


@torch.no_grad()
def detect_images(model: YOLOv8, images: Sequence[np.ndarray], input_size=(640, 640), conf_thresh=0.25, iou_thresh_nms=0.45,
                  class_names: Optional[Sequence[str]] = None, max_det: Optional[int] = None):
    """images: uint8 RGB HWC arrays of arbitrary sizes.  Returns, per image, the list of JSON records the reference
    writes (boxes in ORIGINAL image coordinates, tools/test.py:220-231)."""
    img_h, img_w = input_size
    dev = next(model.parameters()).device
    batch = preprocess_batch(images, (img_h, img_w), device=dev)
    boxes, scores, labels, keep, count = model.detect(batch, conf_thresh, iou_thresh_nms)
    a = boxes.shape[1]
    dets = gather_detections(boxes, scores, labels, keep, count, a if max_det is None else max_det)   # [B, K, 6]
    dets_h = dets.cpu().numpy()
    counts = count.cpu().numpy()
    out = []
    for i, im in enumerate(images):
        oh, ow = im.shape[:2]
        sx = np.float32(ow / img_w)                                   # tools/test.py:222-223 (fp32 tensor *= python float)
        sy = np.float32(oh / img_h)
        recs = []
        n = int(min(counts[i], dets_h.shape[1]))
        for r in range(n):
            x1, y1, x2, y2, sc, lb = dets_h[i, r]
            box = [float(np.float32(x1) * sx), float(np.float32(y1) * sy), float(np.float32(x2) * sx), float(np.float32(y2) * sy)]
            cid = int(lb)
            name = class_names[cid] if class_names is not None and cid < len(class_names) else f"class_{cid}"
            recs.append({"box_xyxy": [round(c, 2) for c in box], "score": round(float(sc), 4), "class_id": cid, "class_name": name})
        out.append(recs)
    return out


@torch.no_grad()
def test(config_path, checkpoint_path, source_path, output_dir="runs/detect/exp", conf_thresh=0.25, iou_thresh_nms=0.45,
         batch_size=32):
    """Same arguments and outputs (JSON files) as the reference's ``test()``; ``batch_size`` is new."""
    from PIL import Image
    cfg = load_config(config_path)
    device = torch.device(cfg.get("device", "cuda"))
    if device.type != "cuda":
        raise RuntimeError("yolo_ms_b200 runs on CUDA devices only (there is no CPU path)")
    os.makedirs(output_dir, exist_ok=True)
    model_cfg, dataset_cfg = cfg["model"], cfg["dataset"]
    class_names = dataset_cfg.get("class_names", [f"class_{i}" for i in range(dataset_cfg["num_classes"])])
    model = YOLOv8(version=model_cfg["architecture"], num_classes=dataset_cfg["num_classes"])
    load_checkpoint(model, checkpoint_path, strict=True)
    model = model.to(device).eval()
    if torch.all(model.head.stride == 0).item():                       # tools/test.py:109-111
        model.head.stride = torch.tensor([8.0, 16.0, 32.0], device=device)
    img_h, img_w = model_cfg.get("input_size", [640, 640])
    image_paths = list_images(source_path)
    if not image_paths:
        print(f"No images found in directory: {source_path}")
        return []
    written = []
    for s in range(0, len(image_paths), batch_size):
        chunk, images = [], []
        for p in image_paths[s:s + batch_size]:
            try:
                images.append(np.asarray(Image.open(p).convert("RGB")))
                chunk.append(p)
            except Exception as e:  # noqa: BLE001  (tools/test.py:155-157 skips unreadable files)
                print(f"Error loading or transforming image {p}: {e}")
        if not images:
            continue
        for p, recs in zip(chunk, detect_images(model, images, (img_h, img_w), conf_thresh, iou_thresh_nms, class_names)):
            base = os.path.splitext(os.path.basename(p))[0]
            path = os.path.join(output_dir, f"{base}_detections.json")
            with open(path, "w") as f:
                json.dump(recs, f, indent=4)
            written.append(path)
            print(f"{p}: {len(recs)} detections -> {path}")
    print("\nTesting finished.")
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description="Run YOLO-MS / YOLOv8 inference on images (B200).")
    ap.add_argument("--config", type=str, default="config/coco_yolov8.yaml")
    ap.add_argument("--checkpoint", type=str, required=True)
    ap.add_argument("--source", type=str, required=True)
    ap.add_argument("--output_dir", type=str, default="runs/detect/exp")
    ap.add_argument("--conf_thresh", type=float, default=0.25)
    ap.add_argument("--iou_thresh_nms", type=float, default=0.45)
    ap.add_argument("--batch_size", type=int, default=32)
    a = ap.parse_args(argv)
    test(a.config, a.checkpoint, a.source, a.output_dir, a.conf_thresh, a.iou_thresh_nms, a.batch_size)


if __name__ == "__main__":
    main()
