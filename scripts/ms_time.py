"""Device time of the MS-Block branch layer in its three fusion modes (csrc/ms_fused.cu) on the layer shapes of the `s`
model at batch 32 (modules.py::MSBlock: first layer of a branch reads the 2c-wide slice [a | b] with repeated pw1 weights, the
others a c-wide tensor): mode 0 = pw1 + depthwise + pw2 as three launches, 1 = pw1 + (depthwise -> pw2), 2 = one kernel.
Each variant is replayed 10x in a CUDA graph.   python scripts/ms_time.py [only_case_index]
`python scripts/ms_time.py prof <case> <mode>` runs ONE mode eagerly three times (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import ops, YmsError

DEV = "cuda"
# (k, c_in of pw1, expanded channels, c_out of pw2, map side, batch)
CASES = [(3, 64, 64, 32, 160, 32), (3, 128, 128, 64, 80, 32), (3, 64, 128, 64, 80, 32), (5, 256, 256, 128, 40, 32),
         (5, 128, 256, 128, 40, 32), (7, 512, 512, 256, 20, 32), (7, 256, 512, 256, 20, 32)]


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


def operands(ci):
    k, c_in, e_ch, c, hw, b = CASES[ci]
    g = torch.Generator().manual_seed(ci)
    bf = torch.bfloat16
    o = dict(k=k, c_in=c_in, e_ch=e_ch, c=c, hw=hw, b=b)
    o["x"] = torch.randn(b, hw, hw, c_in, generator=g).to(DEV).to(bf)
    o["w1"] = (torch.randn(e_ch, c_in, generator=g) / c_in ** 0.5).to(DEV).to(bf)
    o["b1"] = (torch.randn(e_ch, generator=g) * 0.2).to(DEV)
    o["wd"] = (torch.randn(k * k, e_ch, generator=g) / k).to(DEV); o["bd"] = (torch.randn(e_ch, generator=g) * 0.2).to(DEV)
    o["w2"] = (torch.randn(c, e_ch, generator=g) / e_ch ** 0.5).to(DEV).to(bf); o["b2"] = (torch.randn(c, generator=g) * 0.2).to(DEV)
    o["e"] = torch.randn(b, hw, hw, e_ch, generator=g).to(DEV).to(bf); o["d"] = torch.empty_like(o["e"])
    o["y"] = torch.empty(b, hw, hw, c, device=DEV, dtype=bf)
    return o


def plan(o, mode):
    if mode == 0:
        return ops.MsLayerPlan(0, o["d"], o["k"], o["wd"], o["bd"], e=o["e"])
    if mode == 1:
        return ops.MsLayerPlan(1, o["y"], o["k"], o["wd"], o["bd"], e=o["e"], w2=o["w2"], bias2=o["b2"])
    return ops.MsLayerPlan(2, o["y"], o["k"], o["wd"], o["bd"], x=o["x"], w1=o["w1"], bias1=o["b1"], w2=o["w2"], bias2=o["b2"])


if len(sys.argv) > 1 and sys.argv[1] == "prof":
    o = operands(int(sys.argv[2]))
    pl = plan(o, int(sys.argv[3]))
    for _ in range(3):
        pl.run()
    torch.cuda.synchronize()
    print("ok", pl.desc)
    sys.exit(0)

only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for ci, (k, c_in, e_ch, c, hw, b) in enumerate(CASES):
    if only is not None and ci != only:
        continue
    o = operands(ci)
    pw1 = ops.ConvPlan(o["x"], o["w1"].unsqueeze(0).contiguous(), o["b1"], o["e"], ksize=1)
    pw2 = ops.ConvPlan(o["d"], o["w2"].unsqueeze(0).contiguous(), o["b2"], o["y"], ksize=1)
    res = {"pw1": timed([pw1.run]), "dw": timed([plan(o, 0).run]), "pw2": timed([pw2.run]), "dw->pw2": timed([plan(o, 1).run])}
    try:
        res["fused"] = timed([plan(o, 2).run])
    except YmsError:
        res["fused"] = float("nan")
    px = b * hw * hw
    best = min(res["pw1"] + res["dw"] + res["pw2"], res["pw1"] + res["dw->pw2"], res["fused"] if res["fused"] == res["fused"] else 1e9)
    print(f"k={k} {c_in}->{e_ch}->{c} @{hw}x{hw}: " + "  ".join(f"{n} {t:.1f}us" for n, t in res.items()) +
          f"  | best layer {best:.1f}us | dw alone {4.0 * px * e_ch / res['dw'] / 1e3:.0f} GB/s, {2.0 * px * e_ch * k * k / res['dw'] / 1e6:.1f} TFLOP/s")
