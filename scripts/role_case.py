"""Role cycle accounting of ONE 3x3/s1 convolution case and variant (the -DYMS_PROF library):
    YMS_LIB=yolo_ms_b200/libyms_b200_prof.so python scripts/role_case.py c_in c_out hw res variant [batch=32]
Prints, as kcycles averaged over the CTAs that issued MMAs: kernel, prologue, MMA-warp total and its waits (A, B, accumulator-empty,
resident weights), producer waits (A-empty, B-empty), epilogue wait for accumulator-full, items per CTA."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("YMS_LIB", os.path.join(ROOT, "yolo_ms_b200", "libyms_b200_prof.so"))
import torch
from yolo_ms_b200 import ops, _lib

ci, co, hw, res, var = (int(v) for v in sys.argv[1:6])
B = int(sys.argv[6]) if len(sys.argv) > 6 else 32
lib = _lib.load()
lib.yms_debug_set_prof.argtypes = [C.c_void_p]
g = torch.Generator().manual_seed(1)
x = torch.randn(B, hw, hw, ci, generator=g).cuda().to(torch.bfloat16)
w = (torch.randn(9, co, ci, generator=g) / (9 * ci) ** 0.5).cuda().to(torch.bfloat16)
b = torch.randn(co, generator=g).cuda()
y = torch.empty(B, hw, hw, co, device="cuda", dtype=torch.bfloat16)
r = torch.randn(B, hw, hw, co, generator=g).cuda().to(torch.bfloat16) if res else None
pl = ops.ConvPlan(x, w, b, y, ksize=3, stride=1, act=True, residual=r, variant=var)
for _ in range(3):
    pl.run()
torch.cuda.synchronize()
buf = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
assert lib.yms_debug_set_prof(buf.data_ptr()) == 1, "not a -DYMS_PROF build"
pl.run()
torch.cuda.synchronize()
lib.yms_debug_set_prof(None)
t = buf.cpu().double()
act = (t[:, 12] > 0) & (t[:, 0] > 0)
m = t[act].mean(0) / 1e3
print(f"3x3 {ci}->{co} @{hw}x{hw}{' +res' if res else ''} v{var}: ctas {int(act.sum())} items/cta {m[11]*1e3:.1f} kernel {m[12]:.1f} prologue {m[9]:.1f} | "
      f"mma total {m[0]:.1f}: wait A {m[1]:.1f} B {m[2]:.1f} acc-empty {m[3]:.1f} W {m[10]:.1f} busy {m[0]-m[1]-m[2]-m[3]-m[10]:.1f} | "
      f"producer {m[4]:.1f}: wait A-empty {m[5]:.1f} B-empty {m[6]:.1f} | epilogue {m[7]:.1f}: wait acc-full {m[8]:.1f}")
