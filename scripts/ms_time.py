"""Device time of the MS-Block branch layer in its three fusion modes (csrc/ms_fused.cu) on the bench shapes of the `s`
model at batch 32: mode 0 = pw1 + depthwise + pw2 as three launches, 1 = pw1 + (depthwise -> pw2), 2 = one kernel.
Each variant is replayed 10x in a CUDA graph.   python scripts/ms_time.py [only_case_index]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import ops, YmsError

DEV = "cuda"
CASES = [(3, 32, 160, 32, True), (3, 64, 80, 32, True), (3, 64, 80, 32, False), (5, 128, 40, 32, True), (7, 256, 20, 32, True)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None


def timed(fns, reps=10):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


for ci, (k, c, hw, b, two) in enumerate(CASES):
    if only is not None and ci != only:
        continue
    g = torch.Generator().manual_seed(ci)
    e_ch = 2 * c
    bf = torch.bfloat16
    x = torch.randn(b, hw, hw, c, generator=g).to(DEV).to(bf)
    x2 = torch.randn(b, hw, hw, c, generator=g).to(DEV).to(bf) if two else None
    w1 = (torch.randn(e_ch, c * (2 if two else 1), generator=g) / c ** 0.5).to(DEV).to(bf)
    b1 = (torch.randn(e_ch, generator=g) * 0.2).to(DEV)
    wd = (torch.randn(k * k, e_ch, generator=g) / k).to(DEV); bd = (torch.randn(e_ch, generator=g) * 0.2).to(DEV)
    w2 = (torch.randn(c, e_ch, generator=g) / e_ch ** 0.5).to(DEV).to(bf); b2 = (torch.randn(c, generator=g) * 0.2).to(DEV)
    e = torch.empty(b, hw, hw, e_ch, device=DEV, dtype=bf); d = torch.empty_like(e)
    y = torch.empty(b, hw, hw, c, device=DEV, dtype=bf)
    pw1 = ops.ConvPlan(x, w1.unsqueeze(0).contiguous(), b1, e, ksize=1, x2=x2)
    pw2 = ops.ConvPlan(d, w2.unsqueeze(0).contiguous(), b2, y, ksize=1)
    m0 = ops.MsLayerPlan(0, d, k, wd, bd, e=e)
    m1 = ops.MsLayerPlan(1, y, k, wd, bd, e=e, w2=w2, bias2=b2)
    res = {"pw1": timed([pw1.run]), "dw": timed([m0.run]), "pw2": timed([pw2.run]), "dw->pw2": timed([m1.run])}
    try:
        m2 = ops.MsLayerPlan(2, y, k, wd, bd, x=x, x2=x2, w1=w1, bias1=b1, w2=w2, bias2=b2)
        res["fused"] = timed([m2.run])
    except YmsError:
        res["fused"] = float("nan")
    px = b * hw * hw
    print(f"k={k} c={c} @{hw}x{hw} two={two}: " + "  ".join(f"{n} {t:.1f}us" for n, t in res.items()) +
          f"  | dw alone {4.0 * px * e_ch / res['dw'] / 1e3:.0f} GB/s, {2.0 * px * e_ch * k * k / res['dw'] / 1e6:.1f} TFLOP/s")
