"""Device time of the 3x3/s1 convolution variants (yms_conv_params.variant: 1 generic implicit GEMM, 2 / 3 halo kernel with 1 / 2
sub-tiles per item, 5 CTA-pair halo kernel, 6 generic CTA pair, 7 CTA-pair halo kernel with virtual-row tiling) on the bench shapes of the `s` model at batch 32; each variant replayed 10x in a CUDA
graph.   python scripts/conv3_variants.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import ops, YmsError

DEV = "cuda"
CASES = [(32, 32, 160, 0), (32, 32, 160, 1), (64, 64, 80, 0), (64, 64, 80, 1), (128, 128, 40, 0), (128, 128, 40, 1), (256, 256, 20, 0),
         (256, 256, 20, 1), (128, 144, 80, 0), (256, 144, 40, 0), (512, 144, 20, 0), (80, 80, 80, 0), (80, 80, 40, 0), (64, 64, 40, 0)]


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
for ci, co, hw, res in CASES:
    g = torch.Generator().manual_seed(ci + co + hw)
    x = torch.randn(B, hw, hw, ci, generator=g).to(DEV).to(torch.bfloat16)
    w = (torch.randn(9, co, ci, generator=g) / (9 * ci) ** 0.5).to(DEV).to(torch.bfloat16)
    b = torch.randn(co, generator=g).to(DEV)
    y = torch.empty(B, hw, hw, co, device=DEV, dtype=torch.bfloat16)
    r = torch.randn(B, hw, hw, co, generator=g).to(DEV).to(torch.bfloat16) if res else None
    out, ref = [], None
    for v in (1, 2, 3, 5, 6, 7):
        try:
            pl = ops.ConvPlan(x, w, b, y, ksize=3, stride=1, act=True, residual=r, variant=v)
        except YmsError as e:
            out.append(f"v{v} n/a")
            continue
        t = timed(pl.run)
        pl.run(); torch.cuda.synchronize()
        if ref is None:
            ref = y.float().clone()
            d = 0.0
        else:
            d = float((y.float() - ref).norm() / ref.norm())
        out.append(f"v{v} {t:6.1f}us (d {d:.1e})")
    print(f"3x3 {ci:3d}->{co:3d} @{hw}x{hw}{' +res' if res else '     '}: " + "  ".join(out), flush=True)
