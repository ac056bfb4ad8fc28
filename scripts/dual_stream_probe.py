"""Would two half-batches on two streams beat one batch?  Two model instances, batch B/2 each, their one-graph detect() steps
replayed concurrently on two streams, against one instance at batch B.  Kernels of independent graphs fill each other's launch /
drain gaps and the SMs that small-map layers leave idle, but every launch has half the tiles per CTA.
    python scripts/dual_stream_probe.py [version=s] [batch=32] [block=c2f]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from yolo_ms_b200 import YOLOv8, synth

version = sys.argv[1] if len(sys.argv) > 1 else "s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
block = sys.argv[3] if len(sys.argv) > 3 else "c2f"
dev = torch.device("cuda", 0)


def make(b, seed):
    m = YOLOv8(version=version, num_classes=80, block=block)
    m.load_state_dict(synth.calibrated_state_dict(version, block) if hasattr(synth, "calibrated_state_dict") else synth.synthetic_state_dict(m, version, block, seed=1))
    m = m.to(dev).eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    x = synth.make_images(b, 640, 640, seed=seed).to(dev)
    for _ in range(4):
        m.detect(x, 0.25, 0.45)
    torch.cuda.synchronize()
    return m, x


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


m, x = make(B, 7)
t1 = timed(lambda: m.detect(x, 0.25, 0.45))
print(f"one stream, batch {B}: {t1:.4f} ms/step, {B / t1 * 1e3:.0f} images/s", flush=True)
for parts in (2, 4):
    ms = [make(B // parts, 7 + i) for i in range(parts)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]
    main = torch.cuda.current_stream(dev)

    def step():
        ev = torch.cuda.Event(); ev.record(main)
        for (mm, xx), st in zip(ms, streams):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                mm.detect(xx, 0.25, 0.45)
        for st in streams:
            main.wait_stream(st)
    t2 = timed(step)
    print(f"{parts} streams, batch {B // parts} each: {t2:.4f} ms/step, {B / t2 * 1e3:.0f} images/s", flush=True)
    del ms
# two FULL batches in flight: step i + 1 on the other stream, no dependency between them (the NMS of one step, which leaves most SMs
# idle while its longest class chain finishes, then runs next to the other step's convolutions)
ms = [make(B, 7 + i) for i in range(2)]
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
main = torch.cuda.current_stream(dev)


def step2():
    ev = torch.cuda.Event(); ev.record(main)
    for (mm, xx), st in zip(ms, streams):
        st.wait_event(ev)
        with torch.cuda.stream(st):
            mm.detect(xx, 0.25, 0.45)
    for st in streams:
        main.wait_stream(st)


t3 = timed(step2)
print(f"2 streams, batch {B} each (two steps in flight): {t3 / 2:.4f} ms/step, {2 * B / t3 * 1e3:.0f} images/s", flush=True)
