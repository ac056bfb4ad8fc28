"""Per-layer LIBRARY bar on the same GPU: every convolution launch of one forward program next to what the reference's own
module would run for it in eager PyTorch -- `Conv.forward` = nn.Conv2d (cuDNN, bf16 channels_last) -> nn.BatchNorm2d (eval) ->
nn.SiLU (yolov8/model/components.py:69-77), plus the residual add of a Bottleneck (components.py:87-93) where our launch fuses
it.  Both sides are timed the same way (the launch / the three eager ops captured 10x back to back in a CUDA graph: device time,
no host gaps).  cuDNN picks its algorithm with benchmark mode on.  Layers whose input is a folded concat / upsample or whose
epilogue is the fused decode are compared with the plain convolution of the same shape.
    python scripts/cudnn_layers.py [version=s] [batch=32] [hw=640] [out=gpurun_out/cudnn_layers.md]"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from yolo_ms_b200 import YOLOv8, synth

version = sys.argv[1] if len(sys.argv) > 1 else "s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 640
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", "cudnn_layers.md")
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = YOLOv8(version=version, num_classes=80)
model.load_state_dict(synth.synthetic_state_dict(model, version, "c2f", seed=1))
model = model.to(dev).eval()
model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(B, HW, HW, seed=7).to(dev)
model.forward_raw(x)
prog = next(iter(model._programs().values()))[0]
reps = 10


def graph_us(fn):
    fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


rows = []
pat = re.compile(r"conv(\d)x\d/s(\d) (\d+)(?:\+(\d+))?->(\d+) @(\d+)x(\d+)(.*)")
for st, nm in zip(prog.steps, prog.names):
    m = pat.match(nm)
    if not m:
        continue
    k, s, c1, c2, co, h, w, rest = int(m[1]), int(m[2]), int(m[3]), int(m[4] or 0), int(m[5]), int(m[6]), int(m[7]), m[8]
    ci = c1 + c2
    ours = graph_us(st)
    xin = torch.randn(B, ci, h * s, w * s, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(co, ci, k, k, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    rm, rv = torch.zeros(co, device=dev), torch.ones(co, device=dev)
    ga, be = torch.ones(co, device=dev, dtype=torch.bfloat16), torch.zeros(co, device=dev, dtype=torch.bfloat16)
    res = torch.randn(B, co, h, w, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last) if "+res" in rest else None
    linear = ("f32" in rest) or ("decode" in rest)          # the head's final biased convs: no BN / SiLU in the reference either
    conv = lambda: F.conv2d(xin, wt, None, s, k // 2)
    def unit():
        y = F.conv2d(xin, wt, None, s, k // 2)
        if not linear:
            y = F.silu(F.batch_norm(y, rm, rv, ga, be, False, 0.0, 1e-3))
        if res is not None:
            y = y + res
        return y
    t_conv, t_unit = graph_us(conv), graph_us(unit)
    rows.append((nm, ours, t_conv, t_unit))
    del xin, wt, res
tot = [sum(r[i] for r in rows) for i in (1, 2, 3)]
lines = [f"# our conv launches vs the reference's eager modules on the same B200 (version {version}, batch {B}, {HW}x{HW}; device time, us)",
         "", f"total: ours {tot[0]:.1f} | cuDNN convolutions alone {tot[1]:.1f} | Conv.forward eager (conv + BN + SiLU [+ residual]) {tot[2]:.1f}"
             f" -> {tot[2] / tot[0]:.2f}x", "",
         "| launch | ours | cuDNN conv alone | eager Conv.forward | eager / ours |", "|---|---|---|---|---|"]
for nm, a, b, c in rows:
    lines.append(f"| {nm} | {a:.1f} | {b:.1f} | {c:.1f} | {c / a:.2f} |")
os.makedirs(os.path.dirname(out), exist_ok=True)
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
