"""Class-size statistics of the bench workload's NMS input and the kernel time.
    python scripts/nms_probe.py [--ms] [--stats] [name=value ...]     # library options, e.g. nms_groups=8 nms_mask_tiles=100 nms_sort_bitonic=1"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import YOLOv8, synth, ops, _lib
for a in sys.argv[1:]:
    if "=" in a:
        _lib.set_debug_option(a.split("=")[0], int(a.split("=")[1]))
dev = torch.device("cuda", 0)
block = "ms" if "--ms" in sys.argv else "c2f"
model = YOLOv8(version="s", num_classes=80, block=block)
model.load_state_dict(synth.synthetic_state_dict(model, "s", block, seed=1))
model = model.to(dev).eval(); model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(32, 640, 640, seed=7).to(dev)
boxes, scores, labels, keep, count = model.detect(x, 0.25, 0.45)
boxes, scores, labels = boxes.clone(), scores.clone(), labels.clone()
if "--stats" in sys.argv:
    for b in (0, 1, 2, 3):
        valid = scores[b] > 0.25
        h = torch.bincount(labels[b][valid].long(), minlength=80)
        kl = labels[b][keep[b, :int(count[b])].long()].long()
        hk = torch.bincount(kl, minlength=80)
        top = torch.argsort(h, descending=True)[:8]
        print(f"image {b}: candidates {int(valid.sum())}, kept {int(count[b])}, classes used {int((h > 0).sum())}; top classes (n, kept):",
              [(int(h[c]), int(hk[c])) for c in top])
for _ in range(3):
    ops.nms_batched(boxes, scores, labels, 0.25, 0.45, 80)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.nms_batched(boxes, scores, labels, 0.25, 0.45, 80)
b.record(); torch.cuda.synchronize()
print("nms ms:", round(a.elapsed_time(b) / 20, 4), "options:", [a for a in sys.argv[1:] if "=" in a])
