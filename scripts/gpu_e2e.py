"""End-to-end bring-up on the B200 box: CUDA path vs oracle, stage by stage, plus a first timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import weights as W, yolov8_oracle as O, postprocess as PP
from yolo_ms_b200 import YOLOv8, postprocess_batched, launch_count

dev = "cuda"
def rel(a, b): return float((a.float().cpu() - b).norm() / b.norm())

def nchw(t): return t.permute(0, 3, 1, 2)

for version, block, hw, B in () if os.environ.get("E2E_PERF_ONLY") else (("n", "c2f", (64, 96), 2), ("n", "c2f", (320, 320), 2), ("s", "c2f", (640, 640), 2),
                              ("n", "ms", (320, 320), 2), ("s", "ms", (640, 640), 1), ("m", "c2f", (320, 320), 1)):
    sd = W.calibrated_state_dict(version, seed=1, block=block)
    x = W.make_images(B, *hw, seed=7)
    with torch.no_grad():
        ref = O.forward(sd, x, return_parts=True)
        emu = O.forward_bf16_contract(sd, x, return_parts=True)
        # reference-style bf16 noise floor: same oracle with weights+input rounded to bf16 (fp32 math)
    m = YOLOv8(version=version, num_classes=80, block=block)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    xg = x.to(dev)
    t0 = time.time()
    pred = m(xg)
    torch.cuda.synchronize()
    t_first = time.time() - t0
    raws = m.forward_raw(xg)
    taps = m.__dict__["_taps"]
    print(f"{version}/{block} {hw} vs bf16-contract oracle: P", ['%.2e' % rel(nchw(a), b) for a, b in zip(taps['p'], emu['p'])],
          "N", ['%.2e' % rel(nchw(a), b) for a, b in zip(taps['n'], emu['n'])],
          "raw", ['%.2e' % rel(nchw(a), b) for a, b in zip(raws, emu['raw'])], flush=True)
    r_raw = [rel(nchw(a), b) for a, b in zip(raws, ref["raw"])]
    bx = float((pred[..., :4].cpu() - ref["pred"][..., :4]).abs().mean())
    bxm = float((pred[..., :4].cpu() - ref["pred"][..., :4]).abs().max())
    sc = float((pred[..., 4:].cpu() - ref["pred"][..., 4:]).abs().max())
    am = float((pred[..., 4:].argmax(-1).cpu() == ref["pred"][..., 4:].argmax(-1)).float().mean())
    # timing
    for _ in range(3): m(xg)
    torch.cuda.synchronize()
    t0 = time.time(); n = 10
    for _ in range(n): m(xg)
    torch.cuda.synchronize()
    ms = (time.time() - t0) / n * 1e3
    prog = list(m._programs().values())[0][0]
    print(f"{version}/{block} {hw} B={B}: raw relL2={['%.3e' % r for r in r_raw]} box mean/max err px={bx:.3f}/{bxm:.2f} "
          f"score maxerr={sc:.3f} argmax agree={am:.3f} first={t_first:.2f}s step={ms:.3f}ms launches/prog={prog.launches} "
          f"GF={prog.flops/1e9:.2f}", flush=True)
    # post-process parity on OUR pred (identical inputs to both)
    b_, s_, l_, keep, cnt = postprocess_batched(pred, 0.25, 0.45)
    ok = True
    pc = pred.cpu().numpy()
    for i in range(B):
        want, *_ = PP.postprocess_image(pc[i], 0.25, 0.45, PP.greedy_nms_c)
        ok &= bool(np.array_equal(keep[i, :int(cnt[i])].cpu().numpy(), want))
    print(f"   postprocess on CUDA pred exact={ok} kept={cnt.tolist()} total launches so far={launch_count()}", flush=True)

# throughput preview: s, B=32, 640
sd = W.calibrated_state_dict("s", seed=1)
m = YOLOv8(version="s", num_classes=80); m.load_state_dict(sd); m = m.to(dev).eval(); m.head.stride = torch.tensor([8., 16., 32.])
x = W.make_images(32, 640, 640).to(dev)
for _ in range(3): m.detect(x)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10): m.detect(x)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"s B=32 640: detect step {ms:.3f} ms -> {32 / ms * 1e3:.0f} img/s", flush=True)
# per-op timing (eager, events)
prog = list(m._programs().values())[0][0]
times = []
for i, st in enumerate(prog.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st(); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
print("per-step ms (eager):", " ".join(f"{t:.3f}" for t in times), "sum=%.3f" % sum(times))
conv_i = 0
for i, st in enumerate(prog.steps):
    if getattr(st, "__self__", None) is not None and hasattr(st.__self__, "flops"):
        pl = st.__self__
        kp = pl._keep
        xs, ys = tuple(kp[0].shape), tuple(kp[3].shape)
        k = int(round((kp[1].shape[0]) ** 0.5))
        ideal = max(pl.bytes / 6.5e12, pl.flops / 1.4e15) * 1e6
        print(f"  conv {i:2d}: {times[i]*1e3:6.1f} us (ideal {ideal:5.1f}) k{k} {xs[1]}x{xs[2]}x{kp[1].shape[2]}->{ys[1]}x{ys[2]}x{ys[3]} res={kp[4] is not None} {pl.flops/1e9:6.2f} GF {pl.flops/times[i]/1e9:6.1f} TF/s {pl.bytes/1e6:6.1f} MB {pl.bytes/times[i]/1e6:5.0f} GB/s")
