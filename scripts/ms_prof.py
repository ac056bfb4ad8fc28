"""Run ONE fusion mode of the MS-Block layer kernel eagerly (for ncu):  python scripts/ms_prof.py <case> <mode> [reps]
cases as in scripts/ms_time.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import ops

DEV = "cuda"
CASES = [(3, 32, 160, 32, True), (3, 64, 80, 32, True), (3, 64, 80, 32, False), (5, 128, 40, 32, True), (7, 256, 20, 32, True)]
k, c, hw, b, two = CASES[int(sys.argv[1])]
mode = int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator().manual_seed(0)
e_ch, bf = 2 * c, torch.bfloat16
x = torch.randn(b, hw, hw, c, generator=g).to(DEV).to(bf)
x2 = torch.randn(b, hw, hw, c, generator=g).to(DEV).to(bf) if two else None
w1 = (torch.randn(e_ch, c * (2 if two else 1), generator=g) / c ** 0.5).to(DEV).to(bf)
b1 = (torch.randn(e_ch, generator=g) * 0.2).to(DEV)
wd = (torch.randn(k * k, e_ch, generator=g) / k).to(DEV); bd = (torch.randn(e_ch, generator=g) * 0.2).to(DEV)
w2 = (torch.randn(c, e_ch, generator=g) / e_ch ** 0.5).to(DEV).to(bf); b2 = (torch.randn(c, generator=g) * 0.2).to(DEV)
e = torch.randn(b, hw, hw, e_ch, generator=g).to(DEV).to(bf)
y = torch.empty(b, hw, hw, c if mode else e_ch, device=DEV, dtype=bf)
if mode == 0:
    plan = ops.MsLayerPlan(0, y, k, wd, bd, e=e)
elif mode == 1:
    plan = ops.MsLayerPlan(1, y, k, wd, bd, e=e, w2=w2, bias2=b2)
else:
    plan = ops.MsLayerPlan(2, y, k, wd, bd, x=x, x2=x2, w1=w1, bias1=b1, w2=w2, bias2=b2)
for _ in range(reps):
    plan.run()
torch.cuda.synchronize()
print("ok", plan.desc)
