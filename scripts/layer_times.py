"""Per-launch table of one forward program on the GPU: CUDA-event time of every step (eager replay,
mean of N), algorithmic bytes / FLOPs, achieved GB/s and TFLOP/s, and the gap to the roofline
max(bytes / HBM peak, flops / tensor peak).  Usage:
    python scripts/layer_times.py [version=s] [batch=32] [hw=640] [block=c2f] [out=gpurun_out/layers.md]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import YOLOv8, ops
from yolo_ms_b200 import synth

version = sys.argv[1] if len(sys.argv) > 1 else "s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 640
block = sys.argv[4] if len(sys.argv) > 4 else "c2f"
out = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "gpurun_out", f"layers_{version}_{B}_{HW}_{block}.md")
peaks = {"hbm": 6555.8e9, "tc": 1375.4e12}
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peaks = {"hbm": pk["hbm_gbs"] * 1e9, "tc": pk["bf16_tflops_sustained"] * 1e12}
except Exception:
    pass
dev = torch.device("cuda", 0)
model = YOLOv8(version=version, num_classes=80, block=block)
model.load_state_dict(synth.synthetic_state_dict(model, version, block, seed=1))
model = model.to(dev).eval()
model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(B, HW, HW, seed=7).to(dev)
model.forward_raw(x)
prog = next(iter(model._programs().values()))[0]
reps = 10
n = len(prog.steps)
per = [0.0] * n
# device time per launch: every step is captured alone in a CUDA graph, `reps` times back to back (no host launch gaps;
# the eager event pairs used before included the host's launch latency for every kernel shorter than it)
for st in prog.steps:
    st()
torch.cuda.synchronize()
for i, st in enumerate(prog.steps):
    if i < prog.eager_prefix and False:
        continue
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        st()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            st()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    per[i] = a.elapsed_time(b) * 1e3 / reps
rows = []
for i, (t, st, nm) in enumerate(zip(per, prog.steps, prog.names)):
    pl = getattr(st, "__self__", None)
    if pl is not None and hasattr(pl, "flops"):
        fl, by = pl.flops, pl.bytes
    else:
        fl, by = prog.costs.get(i, (0.0, 0.0))
    ideal = max(by / peaks["hbm"], fl / peaks["tc"]) * 1e6
    rows.append((i, nm, t, by / 1e6, fl / 1e9, by / t / 1e3 if t else 0, fl / t / 1e6 if t else 0, ideal, t - ideal))
tot = sum(r[2] for r in rows); tid = sum(r[7] for r in rows)
lines = [f"# per-launch times: version {version}, batch {B}, {HW}x{HW}, block {block} (device time: each launch replayed {reps}x back to back in a CUDA graph)",
         f"total {tot:.1f} us, roofline ideal {tid:.1f} us (HBM {peaks['hbm']/1e9:.0f} GB/s, tensor {peaks['tc']/1e12:.0f} TF/s)", "",
         "| # | step | us | MB | GFLOP | GB/s | TF/s | ideal us | gap us |", "|---|---|---|---|---|---|---|---|---|"]
for r in rows:
    lines.append(f"| {r[0]} | {r[1]} | {r[2]:.1f} | {r[3]:.1f} | {r[4]:.1f} | {r[5]:.0f} | {r[6]:.0f} | {r[7]:.1f} | {r[8]:.1f} |")
os.makedirs(os.path.dirname(out), exist_ok=True)
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
