"""CPU: the C-ABI library loads and exports every symbol include/yms_b200.h declares; the
host-side mirrors keep the reference's state_dict layout; nothing falls back to the CPU."""
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from yolo_ms_b200 import _lib
    header = open(os.path.join(ROOT, "include", "yms_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(yms_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.yms_abi_version() == 1
    assert _lib.launch_count() >= 0


def test_argument_errors_are_reported_without_a_gpu():
    import ctypes as C
    from yolo_ms_b200 import _lib
    lib = _lib.load()
    p = _lib.ConvParams()
    h = C.c_void_p()
    assert lib.yms_conv_plan_create(C.byref(p), C.byref(h)) == -1          # YMS_E_ARG
    assert b"bad sizes" in lib.yms_last_error()
    assert lib.yms_nms_workspace_bytes(0, 8400) == 0
    assert 0 < lib.yms_nms_workspace_bytes(2, 8400) < lib.yms_nms_workspace_bytes(2, 30000)
    assert lib.yms_nms_batched(None, None, None, None, 1, 10, 5000, 0.25, 0.45, None, None, None, 0, None) == -2


def test_plan_modifiers_report_argument_errors_without_a_gpu():
    """yms_conv_plan_fuse_decode / yms_conv_plan_add_upsampled validate their arguments before touching the device."""
    import ctypes as C
    from yolo_ms_b200 import _lib
    lib = _lib.load()
    f = _lib.DecodeFusion()
    assert lib.yms_conv_plan_fuse_decode(None, C.byref(f)) == -1 and b"null" in lib.yms_last_error()
    assert lib.yms_conv_plan_add_upsampled(None, None, 64, 8, 8) == -1


def test_host_side_fusion_switches():
    """Which programs decode in the conv epilogue is host logic: class counts up to 128 after the host has zero-padded the class
    branch to a multiple of 16 (the fused epilogue's limits, include/yms_b200.h; the reference's own configs use 1 and 10
    classes); everything else keeps the stand-alone decode kernel.  The autotuner's cache key separates layers that differ only
    in buffer strides (a channel slice of a concat buffer vs a dense tensor)."""
    from yolo_ms_b200 import engine
    from yolo_ms_b200.model.yolov8_head import Head
    assert Head(version="n", num_classes=80).can_fuse_decode()
    assert Head(version="n", num_classes=16).can_fuse_decode()
    for nc, ncp in ((1, 16), (10, 16), (24, 32), (80, 80), (128, 128)):
        h = Head(version="n", num_classes=nc)
        assert h.nc_pad == ncp and h.no == 64 + nc and h.can_fuse_decode()
        assert h.state_dict()["cls.0.2.weight"].shape[0] == nc            # the state_dict keeps the reference's shapes
    assert not Head(version="n", num_classes=144).can_fuse_decode()
    assert not Head(version="n", num_classes=80, ch=8).can_fuse_decode()
    dense = torch.empty(2, 8, 8, 32)
    sliced = torch.empty(2, 8, 8, 64)[..., :32]
    y = torch.empty(2, 8, 8, 32)
    k1 = engine._tune_key("3x3", dense, y, 3, 1, True, None)
    assert k1 == engine._tune_key("3x3", torch.empty(2, 8, 8, 32), torch.empty(2, 8, 8, 32), 3, 1, True, None)
    assert k1 != engine._tune_key("3x3", sliced, y, 3, 1, True, None)
    assert k1 != engine._tune_key("3x3", dense, y, 3, 1, True, y)
    assert k1 != engine._tune_key("s2pair", dense, y, 3, 1, True, None)


def test_cpu_tensors_raise_not_fall_back():
    from yolo_ms_b200 import YmsError, ops, postprocess
    from yolo_ms_b200.yolov8 import YOLOv8
    with pytest.raises(YmsError):
        ops.nms_batched(torch.zeros(1, 4, 4), torch.zeros(1, 4), torch.zeros(1, 4, dtype=torch.int32), 0.25, 0.45, 80)
    with pytest.raises(YmsError):
        postprocess(torch.zeros(1, 10, 84))
    with pytest.raises(YmsError):
        YOLOv8(version="n", num_classes=80)(torch.zeros(1, 3, 64, 64))


@pytest.mark.parametrize("version", ["n", "s", "m"])
def test_state_dict_layout_matches_reference_manifest(version):
    from oracle import weights as W
    from yolo_ms_b200.yolov8 import YOLOv8
    m = YOLOv8(version=version, num_classes=80)
    man = W.load_manifest(version)
    sd = m.state_dict()
    assert list(sd.keys()) == list(man.keys())
    assert all(list(sd[k].shape) == man[k] for k in man)
    m.load_state_dict(W.make_state_dict(man), strict=True)
    assert isinstance(m.head.stride, torch.Tensor) and "stride" not in " ".join(sd.keys())
    # {'model': ...} / 'module.' wrappers are handled by the caller in the reference (tools/utils.py:55-67);
    # the module itself takes a plain state_dict.
    ms = YOLOv8(version=version, num_classes=80, block="ms")
    assert set(ms.state_dict().keys()) == set(W.load_manifest(version, "ms").keys())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yolo_ms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
