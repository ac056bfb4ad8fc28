"""Probe: how much does capturing stem + decode + NMS into the step's CUDA graph save vs the current detect()?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import YOLOv8, synth, ops
dev = torch.device("cuda", 0)
model = YOLOv8(version="s", num_classes=80); model.load_state_dict(synth.synthetic_state_dict(model, "s", "c2f", seed=1))
model = model.to(dev).eval(); model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(32, 640, 640, seed=7).to(dev)
for _ in range(3): out = model.detect(x, 0.25, 0.45)
torch.cuda.synchronize()
def timeit(fn, n=30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn()
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print("detect() now: %.4f ms" % timeit(lambda: model.detect(x, 0.25, 0.45)))
prog, io = next(iter(model._programs().values()))
def full():
    for s in prog.steps[:prog.eager_prefix]: s()
    prog.run_eager(start=prog.eager_prefix)
    pred, (cb, cs, cl) = ops.head_decode(io["outputs"], [8.0, 16.0, 32.0], 80, with_candidates=True)
    return ops.nms_batched(cb, cs, cl, 0.25, 0.45, 80)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    full(); full()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    res = full()
torch.cuda.synchronize()
print("one graph for the whole step: %.4f ms" % timeit(g.replay))
