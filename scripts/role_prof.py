"""Who waits for whom inside the tcgen05 conv kernels: runs every conv launch of one forward program with the
-DYMS_PROF build (python -m yolo_ms_b200.build --prof) and prints, per launch, the cycles each warp role spent
blocked on its mbarriers (mean over CTAs).  Usage (GPU box):
    YMS_LIB=yolo_ms_b200/libyms_b200_prof.so python scripts/role_prof.py [version] [batch] [hw] [block]
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("YMS_LIB", os.path.join(ROOT, "yolo_ms_b200", "libyms_b200_prof.so"))
import torch
from yolo_ms_b200 import YOLOv8, synth, _lib

version = sys.argv[1] if len(sys.argv) > 1 else "s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 640
block = sys.argv[4] if len(sys.argv) > 4 else "c2f"
out = os.path.join(ROOT, "gpurun_out", f"roles_{version}_{B}_{HW}_{block}.md")
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.yms_debug_set_prof.argtypes = [C.c_void_p]
model = YOLOv8(version=version, num_classes=80, block=block)
model.load_state_dict(synth.synthetic_state_dict(model, version, block, seed=1))
model = model.to(dev).eval()
model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(B, HW, HW, seed=7).to(dev)
model.forward_raw(x)
prog = next(iter(model._programs().values()))[0]
buf = torch.zeros(148, 16, dtype=torch.int64, device=dev)
assert lib.yms_debug_set_prof(buf.data_ptr()) == 1, "not a -DYMS_PROF build"
for st in prog.steps: st()
torch.cuda.synchronize()
hdr = "| # | step | ctas | tiles/cta | kernel kcyc | prologue (barriers / tmem alloc / sync) | mma total | mma: wait A | wait B | wait acc-empty | wait W | mma busy | prod wait A-empty | wait B-empty | epi wait acc-full |"
lines = [f"# role cycle accounting: version {version}, batch {B}, {HW}x{HW}, block {block} (kcycles, mean over CTAs)", "", hdr, "|" + "---|" * 15]
for i, (st, nm) in enumerate(zip(prog.steps, prog.names)):
    if not nm.startswith("conv"):
        continue
    buf.zero_()
    st()
    torch.cuda.synchronize()
    b = buf.cpu().double()
    act = (b[:, 12] > 0) & (b[:, 0] > 0)               # CTAs that issued MMAs (the leaders of CTA pairs)
    n = int(act.sum())
    if n == 0:
        continue
    m = b[act].mean(0) / 1e3
    busy = m[0] - m[1] - m[2] - m[3] - m[10]
    lines.append(f"| {i} | {nm} | {n} | {m[11]*1e3:.1f} | {m[12]:.1f} | {m[9]:.1f} ({m[15]:.1f} / {m[14]:.1f} / {m[13]:.1f}) | {m[0]:.1f} | {m[1]:.1f} | {m[2]:.1f} | {m[3]:.1f} | {m[10]:.1f} | {busy:.1f} | {m[5]:.1f} | {m[6]:.1f} | {m[8]:.1f} |")
lib.yms_debug_set_prof(None)
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
