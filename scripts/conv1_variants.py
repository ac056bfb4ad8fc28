"""Device time of the generic convolution kernel, single CTA (variant 0) vs CTA pair (variant 5), on the 1x1 and 3x3/s2 shapes of
the `s` model at batch 32; each variant replayed 10x in a CUDA graph.   python scripts/conv1_variants.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import ops, YmsError

DEV = "cuda"
# (k, stride, c_in, c_out, input map)
CASES = [(1, 1, 64, 64, 160), (1, 1, 96, 64, 160), (1, 1, 128, 128, 80), (1, 1, 256, 128, 80), (1, 1, 192, 128, 80), (1, 1, 256, 256, 40),
         (1, 1, 512, 256, 40), (1, 1, 384, 256, 40), (1, 1, 512, 512, 20), (1, 1, 768, 512, 20), (1, 1, 512, 256, 20), (1, 1, 1024, 512, 20),
         (3, 2, 64, 128, 160), (3, 2, 128, 256, 80), (3, 2, 256, 512, 40), (3, 2, 128, 128, 80), (3, 2, 256, 256, 40)]


def timed(fn, reps=10):
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


B = 32
for k, s, ci, co, hw in CASES:
    g = torch.Generator().manual_seed(ci + co + hw)
    x = torch.randn(B, hw, hw, ci, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(k * k, co, ci, generator=g) / (k * k * ci) ** 0.5).to(DEV).to(torch.bfloat16)
    b = torch.randn(co, generator=g).to(DEV)
    y = torch.empty(B, hw // s, hw // s, co, device=DEV, dtype=torch.bfloat16)
    out, ref = [], None
    for v in (0, 5):
        try:
            pl = ops.ConvPlan(x, w, b, y, ksize=k, stride=s, act=True, variant=v)
        except YmsError as e:
            out.append(f"v{v} n/a")
            continue
        t = timed(pl.run)
        pl.run(); torch.cuda.synchronize()
        if ref is None:
            ref, d = y.float().clone(), 0.0
        else:
            d = float((y.float() - ref).norm() / ref.norm())
        out.append(f"v{v} {t:6.1f}us (d {d:.1e})")
    print(f"{k}x{k}/s{s} {ci:4d}->{co:3d} @{hw // s}x{hw // s}: " + "  ".join(out), flush=True)
