"""Phase timing of nms_kernel on the bench workload (needs the -DYMS_PROF build)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("YMS_LIB", os.path.join(ROOT, "yolo_ms_b200", "libyms_b200_prof.so"))
import torch
from yolo_ms_b200 import YOLOv8, synth, _lib, ops
dev = torch.device("cuda", 0)
lib = _lib.load(); lib.yms_debug_set_prof.argtypes = [C.c_void_p]
block = "ms" if "--ms" in sys.argv else "c2f"
model = YOLOv8(version="s", num_classes=80, block=block)
model.load_state_dict(synth.synthetic_state_dict(model, "s", block, seed=1))
model = model.to(dev).eval(); model.head.stride = torch.tensor([8.0, 16.0, 32.0])
x = synth.make_images(32, 640, 640, seed=7).to(dev)
raws = model.forward_raw(x)
pred, (cb, cs, cl) = ops.head_decode(raws, [8.0, 16.0, 32.0], 80, with_candidates=True)
for _ in range(2): ops.nms_batched(cb, cs, cl, 0.25, 0.45, 80)
buf = torch.zeros(148 * 16 + 512, dtype=torch.int64, device=dev)
assert lib.yms_debug_set_prof(buf.data_ptr()) == 1
keep, cnt = ops.nms_batched(cb, cs, cl, 0.25, 0.45, 80)
torch.cuda.synchronize(); lib.yms_debug_set_prof(None)
tr = buf.cpu()[128 * 16:128 * 16 + 512].view(128, 4)
b = buf.cpu()[:148 * 16].view(148, 16)[:128].double()
names = ["histogram+ranges", "compact", "sort", "segments+boxes", "suppression (stamp 4->6)", "-", "keep list", "ticket+concat"]
d = b[:, 1:9] - b[:, 0:8]
d[:, 4] = b[:, 6] - b[:, 4]
print("kept per image (first 4):", cnt[:4].tolist())
for i, nm in enumerate(names):
    col = d[:, i][b[:, i + 1] > 0]
    if nm != "-" and len(col): print(f"{nm:24s} mean {col.mean()/1e3:8.1f} kcyc  max {col.max()/1e3:8.1f} kcyc")
srt = b[:, 12] - b[:, 2]; print("sort: warp sorts of segments / runs mean %.1f max %.1f kcyc, rank merges mean %.1f max %.1f kcyc" % (srt.mean()/1e3, srt.max()/1e3, (b[:, 3] - b[:, 12]).mean()/1e3, (b[:, 3] - b[:, 12]).max()/1e3))
tot = (b[:, 7] - b[:, 0]); print("per-CTA total kcyc (first 16 CTAs = 4 images):", [round(float(v) / 1e3) for v in tot[:16]]); print("total to stamp7: mean %.1f max %.1f kcyc" % (tot.mean()/1e3, tot.max()/1e3))

# chunk trace of the largest class of CTA 0 (profiling build): start of the chunk's scan, predecessors final, own result published
t0 = int(b[0, 0])
print("chunk: start  pred_final  done  (kcyc from kernel start) | scan+wait  critical  kept_before")
for j in range(128):
    if int(tr[j, 2]) == 0: break
    s_, p_, d_, k_ = (int(v) for v in tr[j])
    if j < 12 or j % 8 == 0: print(f"{j:4d}: {(s_-t0)/1e3:8.1f} {(p_-t0)/1e3:8.1f} {(d_-t0)/1e3:8.1f} | {(p_-s_)/1e3:7.1f} {(d_-p_)/1e3:7.1f} {k_:5d}")
