import torch, sys
sys.path.insert(0, '.')
from yolo_ms_b200 import ops
dev='cuda'
def t(fn):
    for _ in range(3): fn()
    a=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize(); return a.elapsed_time(e)/20*1e3
for (k,c,hw) in [(3,64,160),(3,128,80),(5,256,40),(7,512,20),(9,64,160),(5,64,160),(7,64,160)]:
    x=torch.randn(32,hw,hw,c,device=dev).to(torch.bfloat16); y=torch.empty_like(x)
    w=torch.randn(k*k,c,device=dev)*0.1; b=torch.randn(c,device=dev)*0.1
    us=t(lambda: ops.dwconv(x,w,b,y,k)); by=2*x.numel()*2
    print(f"dwconv k={k} c={c} @{hw}: {us:.1f} us  {by/us/1e3:.0f} GB/s  {2*x.numel()*k*k/us/1e6:.1f} TFLOP/s")
