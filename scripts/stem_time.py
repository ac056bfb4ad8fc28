import torch, sys
sys.path.insert(0, '.')
from yolo_ms_b200 import ops
dev='cuda'
B=32
x=torch.randn(B,3,640,640,device=dev); xu=torch.randint(0,256,(B,640,640,3),dtype=torch.uint8,device=dev)
w=torch.randn(32,3,3,3,device=dev)*0.2; b=torch.randn(32,device=dev)*0.1
y=torch.empty(B,320,320,32,device=dev,dtype=torch.bfloat16)
def t(fn):
    for _ in range(3): fn()
    a=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    e.record(); torch.cuda.synchronize(); return a.elapsed_time(e)/20*1e3
print("stem f32 us", t(lambda: ops.stem_conv(x,w,b,y)), " u8 us", t(lambda: ops.stem_conv_u8(xu,w,b,y)))
