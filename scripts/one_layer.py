"""Run one conv layer a few times (for ncu / quick timing).  usage: one_layer.py B H W cin cout k stride [res] [f32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolo_ms_b200 import ops
B, H, W, cin, cout, k, s = map(int, sys.argv[1:8])
res = len(sys.argv) > 8 and sys.argv[8] == "1"
f32 = len(sys.argv) > 9 and sys.argv[9] == "1"
dev = "cuda"
x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
w = (torch.randn(k * k, cout, cin, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(cout, device=dev)
y = torch.empty(B, H // s, W // s, cout, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
r = torch.randn(B, H // s, W // s, cout, device=dev).to(torch.bfloat16) if res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
plan = ops.ConvPlan(x, w, b, y, ksize=k, stride=s, act=True, residual=r)
ts = []
for i in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"layer B{B} {H}x{W} {cin}->{cout} k{k} s{s} res={res}: us {['%.1f' % t for t in ts]}  bytes {plan.bytes/1e6:.1f} MB -> {plan.bytes/min(ts)/1e3:.0f} GB/s, {plan.flops/min(ts)/1e6:.0f} TF/s")
