"""bf16 noise of the MS-Block variant end to end: raw head logits / feature taps of the CUDA path vs the CPU oracle under the
same numeric contract, for the shapes tests/test_gpu_model.py gates (sets the gates at 1.5x the measured values)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import weights as W
from oracle import yolov8_oracle as O
from yolo_ms_b200.yolov8 import YOLOv8

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())

for version, hw, batch in (("n", (64, 64), 2), ("n", (256, 256), 2), ("s", (256, 256), 2), ("s", (640, 640), 1)):
    sd = W.calibrated_state_dict(version, seed=1, block="ms")
    m = YOLOv8(version=version, num_classes=80, block="ms")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.head.stride = torch.tensor([8.0, 16.0, 32.0])
    x = W.make_images(batch, *hw, seed=7)
    with torch.no_grad():
        ref = O.forward(sd, x, return_parts=True)
        emu = O.forward_bf16_contract(sd, x, return_parts=True)
    raws = m.forward_raw(x.cuda())
    taps = m.__dict__["_taps"]
    n = lambda t: t.permute(0, 3, 1, 2)
    print(version, hw, "raw vs contract", [round(rel(n(raws[i]), emu["raw"][i]), 4) for i in range(3)],
          "p", [round(rel(n(taps["p"][i]), emu["p"][i]), 4) for i in range(3)],
          "n", [round(rel(n(taps["n"][i]), emu["n"][i]), 4) for i in range(3)],
          "| contract vs fp32 (floor)", [round(rel(emu["raw"][i], ref["raw"][i]), 4) for i in range(3)],
          "| gpu vs fp32", [round(rel(n(raws[i]), ref["raw"][i]), 4) for i in range(3)])
