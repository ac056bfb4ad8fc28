"""NMS input statistics + kernel time of one bench leg:  python scripts/nms_case.py <version> <block> <hw> <batch> [option=value ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import YOLOv8, synth, ops, _lib
version, block, hw, batch = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
for a in sys.argv[5:]:
    if "=" in a:
        _lib.set_debug_option(a.split("=")[0], int(a.split("=")[1]))
dev = torch.device("cuda", 0)
model = YOLOv8(version=version, num_classes=80, block=block)
model.load_state_dict(synth.synthetic_state_dict(model, version, block, seed=1))
model = model.to(dev).eval(); model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(batch, hw, hw, seed=7).to(dev)
boxes, scores, labels, keep, count = [t.clone() for t in model.detect(x, 0.25, 0.45)]
for b in (0, batch - 1):
    valid = scores[b] > 0.25
    h = torch.bincount(labels[b][valid].long(), minlength=80)
    kl = labels[b][keep[b, :int(count[b])].long()].long()
    hk = torch.bincount(kl, minlength=80)
    top = torch.argsort(h, descending=True)[:6]
    print(f"image {b}: candidates {int(valid.sum())} of {scores.shape[1]}, kept {int(count[b])}, classes used {int((h > 0).sum())}; top classes (n, kept):",
          [(int(h[c]), int(hk[c])) for c in top])
pb = ops.PostBuffers(batch, scores.shape[1], dev)
for _ in range(3):
    ops.nms_batched(boxes, scores, labels, 0.25, 0.45, 80, out=pb)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.nms_batched(boxes, scores, labels, 0.25, 0.45, 80, out=pb)
b.record(); torch.cuda.synchronize()
print("nms ms:", round(a.elapsed_time(b) / 10, 4), "options:", [a for a in sys.argv[5:] if "=" in a])
