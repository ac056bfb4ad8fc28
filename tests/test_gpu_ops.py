"""-m gpu: every kernel through the C ABI against its reference, on the B200 box."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from gpu_util import DEV, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from yolo_ms_b200 import ops as _ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


# name, B, H, W, cin, cout, k, stride, act, res, cin2, out_f32, slices
CONV_CASES = [
    ("1x1_64_64", 2, 40, 40, 64, 64, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_tailM", 1, 20, 20, 64, 64, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_16_16", 2, 16, 16, 16, 16, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_48_96", 2, 16, 16, 48, 96, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_80_80_f32", 2, 20, 20, 80, 80, 1, 1, 0, 0, 0, 1, 1),
    ("1x1_768_512", 2, 20, 20, 768, 512, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_576_288", 1, 20, 20, 576, 288, 1, 1, 1, 0, 0, 0, 1),       # several N tiles, c_out % 64 != 0
    ("1x1_two_src_odd", 2, 20, 20, 48, 80, 1, 1, 1, 0, 32, 0, 1),
    ("3x3_64_64_res", 2, 40, 40, 64, 64, 3, 1, 1, 1, 0, 0, 0),
    ("3x3_288_288_res", 1, 20, 20, 288, 288, 3, 1, 1, 1, 0, 0, 1),
    ("3x3_128_80_w20", 2, 20, 20, 128, 80, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_32_32_w160", 1, 160, 160, 32, 32, 3, 1, 1, 0, 0, 0, 0),
    ("3x3s2_64_128", 2, 80, 80, 64, 128, 3, 2, 1, 0, 0, 0, 0),
    ("3x3s2_384_576", 1, 20, 20, 384, 576, 3, 2, 1, 0, 0, 0, 0),
    ("3x3s2_slices", 2, 40, 40, 128, 128, 3, 2, 1, 0, 0, 0, 1),
    ("3x3s2_odd_out", 1, 24, 40, 64, 64, 3, 2, 1, 0, 0, 0, 0),
    ("3x3_f32_320", 1, 64, 64, 64, 64, 3, 1, 0, 0, 0, 1, 0),
    ("3x3_128_128_w40", 4, 40, 40, 128, 128, 3, 1, 1, 1, 0, 0, 1),      # streamed weights, residual, slices, many items per CTA / cluster
    ("3x3_256_256_w20", 8, 20, 20, 256, 256, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_64_144_w80", 2, 80, 80, 64, 144, 3, 1, 1, 0, 0, 0, 0),        # N = 144: three 64-channel output chunks, the last one partial
    ("3x3_h16", 3, 16, 16, 64, 64, 3, 1, 1, 0, 0, 0, 0),                # variant 7 at its smallest map: vh = 18, a band's halo wraps exactly once
    ("1x1_64_1152", 1, 20, 20, 64, 1152, 1, 1, 1, 0, 0, 0, 0),          # more bias values than two per epilogue thread (staging loop)
    ("3x3_h18_res", 3, 18, 24, 64, 64, 3, 1, 1, 1, 0, 0, 0),            # variant 7: bands of 16 virtual rows over images 20 rows apart, the last band runs past the batch
]


def _variants(case):
    """3x3/s1 bf16 single-source layers have four tcgen05 implementations (the per-layer autotuner picks one); variant 5 = the
    CTA-pair kernel (cta_group::2), for c_out <= 256."""
    name, B, H, W, cin, cout, k, s, act, res, cin2, f32, sl = case
    if f32:
        return (0,)
    if not (k == 3 and s == 1 and not cin2):
        return (0, 5)                                   # generic kernel: single CTA / CTA pair
    return (0, 1, 2, 3, 5, 6, 7) if cout <= 256 else (0, 1, 2, 3, 6)       # 7: pair kernel, virtual-row tiling


@pytest.mark.parametrize("case,variant", [(c, v) for c in CONV_CASES for v in _variants(c)],
                         ids=[f"{c[0]}-v{v}" for c in CONV_CASES for v in _variants(c)])
def test_conv_gemm_matches_torch_fp32(ops, case, variant):
    """tcgen05 implicit-GEMM conv vs a plain PyTorch fp32 conv of the same (bf16-rounded) operands.
    Tolerance: bf16 output rounding (2^-9) -> rel-L2 <= 4e-3; fp32 outputs <= 1e-5."""
    name, B, H, W, cin, cout, k, s, act, res, cin2, f32, sl = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    pad_c = 24 if sl else 0

    def mk(c, h, w):
        full = torch.randn(B, h, w, c + pad_c, generator=g).to(DEV).to(torch.bfloat16)
        return full[..., 8:8 + c] if sl else full

    x = mk(cin, H, W)
    x2 = mk(cin2, H, W) if cin2 else None
    ktot = cin + cin2
    wt = (torch.randn(cout, ktot, k, k, generator=g) / (ktot * k * k) ** 0.5).to(DEV).to(torch.bfloat16)
    bias = (torch.randn(cout, generator=g) * 0.5).to(DEV)
    Ho, Wo = H // s, W // s
    r = mk(cout, Ho, Wo) if res else None
    yfull = torch.full((B, Ho, Wo, cout + pad_c), 7.0, device=DEV, dtype=torch.float32 if f32 else torch.bfloat16)
    y = yfull[..., 8:8 + cout] if sl else yfull
    wpk = wt.permute(2, 3, 0, 1).reshape(k * k, cout, ktot).contiguous()
    try:
        ops.ConvPlan(x, wpk, bias, y, ksize=k, stride=s, act=bool(act), residual=r, x2=x2, variant=variant).run()
    except Exception as e:                                # the CTA-pair variants need two M tiles / sub-tiles: "unsupported" is the contract
        if variant in (5, 6, 7) and "code -2" in str(e):
            pytest.skip(str(e))
        raise
    xin = x.float() if x2 is None else torch.cat([x.float(), x2.float()], -1)
    ref = F.conv2d(xin.permute(0, 3, 1, 2), wt.float(), bias, stride=s, padding=k // 2)
    ref = F.silu(ref) if act else ref
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    ref = ref.permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < (1e-5 if f32 else 4e-3)
    if sl:   # neighbours of the channel slice must be untouched (concat-by-pointer contract)
        assert bool((yfull[..., :8] == 7).all() and (yfull[..., 8 + cout:] == 7).all())


@pytest.mark.parametrize("B,H,W,c_low,c_skip,cout", [(2, 40, 40, 512, 256, 256), (3, 20, 28, 256, 128, 128), (1, 80, 80, 64, 32, 32),
                                                     (2, 10, 6, 576, 384, 384),
                                                     (12, 96, 96, 32, 32, 144)])    # several tiles per epilogue group (partial sums prefetched across tiles), partial last chunk
def test_conv1x1_over_upsampled_concat(ops, B, H, W, c_low, c_skip, cout):
    """Neck: conv1x1(cat[upsample2x(a), b]) computed as upsample2x(W_a . a) + W_b . b (yms_conv_plan_add_upsampled): a linear fp32
    1x1 plan at half resolution, then the plan over b adds it before bias + SiLU.  Reference: plain PyTorch fp32 on the same
    bf16-rounded operands (F.interpolate nearest, cat, conv2d); tolerance = bf16 output rounding, rel-L2 <= 4e-3.  Map sizes whose
    128-pixel tiles straddle rows and images are included; y is a channel slice of a wider buffer."""
    g = torch.Generator().manual_seed(c_low + H)
    a = torch.randn(B, H // 2, W // 2, c_low, generator=g).to(DEV).to(torch.bfloat16)
    cat = torch.randn(B, H, W, c_low + c_skip, generator=g).to(DEV).to(torch.bfloat16)       # first slot deliberately garbage
    b = cat[..., c_low:]
    ktot = c_low + c_skip
    wt = (torch.randn(cout, ktot, generator=g) / ktot ** 0.5).to(DEV).to(torch.bfloat16)
    bias = (torch.randn(cout, generator=g) * 0.5).to(DEV)
    part = torch.empty(B, H // 2, W // 2, cout, device=DEV, dtype=torch.float32)
    yfull = torch.full((B, H, W, cout + 16), 7.0, device=DEV, dtype=torch.bfloat16)
    y = yfull[..., 8:8 + cout]
    ops.ConvPlan(a, wt[:, :c_low].contiguous().view(1, cout, c_low), torch.zeros_like(bias), part, ksize=1, act=False).run()
    main = ops.ConvPlan(b, wt[:, c_low:].contiguous().view(1, cout, c_skip), bias, y, ksize=1, act=True)
    main.add_upsampled(part)
    main.run()
    up = F.interpolate(a.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    xin = torch.cat([up, b.float().permute(0, 3, 1, 2)], 1)
    want = F.silu(F.conv2d(xin, wt.float().view(cout, ktot, 1, 1), bias)).permute(0, 2, 3, 1)
    assert rel_l2(y.float(), want) <= 4e-3
    assert float(yfull[..., :8].float().min()) == 7.0 and float(yfull[..., 8 + cout:].float().max()) == 7.0
    with pytest.raises(ops.YmsError):
        main.add_upsampled(part[:, :, :-1])


def test_half_cta_mode_in_a_subprocess():
    """Library option conv_half=1 (opt-in): conv_gemm_kernel with 2 epilogue groups / 320 threads / 256 TMEM
    columns / <= 113 KB so that two CTAs share an SM.  The option is process-wide, so the generic-kernel parity cases and the
    decode-fused program are re-run in a child interpreter that sets it at session start (tests/conftest.py)."""
    import subprocess
    import sys
    env = dict(os.environ, YMS_TEST_OPTIONS="conv_half=1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(here, "test_gpu_ops.py"), os.path.join(here, "test_gpu_model.py"),
                        "-k", "(test_conv_gemm_matches_torch_fp32 and (1x1 or s2)) or test_conv1x1_over_upsampled_concat or "
                              "test_fused_decode_is_bit_identical"],
                       env=env, cwd=os.path.dirname(here), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout


def test_fuse_decode_and_add_upsampled_reject_unsupported_plans(ops):
    """yms_conv_plan_fuse_decode only takes the linear f32-output 1x1 conv that ends a head branch (c_out 64 for the box branch,
    num_classes -- a multiple of 16, <= 128 -- for the class branch); yms_conv_plan_add_upsampled only bf16-output 1x1 convs with
    c_out % 16 == 0 on even maps.  Everything else is an error, never a silent fallback."""
    B, H, W = 1, 8, 12
    x = torch.zeros(B, H, W, 64, device=DEV, dtype=torch.bfloat16)
    w1 = torch.zeros(1, 64, 64, device=DEV, dtype=torch.bfloat16)
    bias = torch.zeros(64, device=DEV)
    pred = torch.zeros(B, H * W, 4 + 80, device=DEV)
    stride = torch.ones(1, device=DEV)
    boxes = torch.zeros(B, H * W, 4, device=DEV)
    yf = torch.zeros(B, H, W, 64, device=DEV)
    yb = torch.zeros(B, H, W, 64, device=DEV, dtype=torch.bfloat16)
    ok = ops.ConvPlan(x, w1, bias, yf, ksize=1, act=False)
    ok.fuse_decode("box", stride, pred, 0, cand_boxes=boxes)                           # the supported case
    ok.run()
    torch.cuda.synchronize()
    assert float(pred[..., 2:4].min()) == 15.0 and float(pred[..., 2:4].max()) == 15.0 and float(pred[0, 13, 0]) == (1 + 0.5) and float(pred[0, 13, 1]) == (1 + 0.5)   # zero logits: l=t=r=b=7.5
    for plan, args in [
        (ops.ConvPlan(x, w1, bias, yb, ksize=1, act=False), ("box", stride, pred, 0)),                       # bf16 output
        (ops.ConvPlan(x, w1, bias, yf, ksize=1, act=True), ("box", stride, pred, 0)),                        # activation
        (ops.ConvPlan(x, torch.zeros(9, 64, 64, device=DEV, dtype=torch.bfloat16), bias, yf, ksize=3, act=False), ("box", stride, pred, 0)),
        (ops.ConvPlan(x, w1, bias, yf, ksize=1, act=False), ("cls", stride, pred, 0)),                       # c_out 64 != 80 classes
        (ops.ConvPlan(x, w1, bias, yf, ksize=1, act=False), ("box", stride, torch.zeros(B, H * W, 4 + 24, device=DEV), 0)),   # 24 classes
        (ops.ConvPlan(x, w1, bias, yf, ksize=1, act=False), ("box", stride, pred, 8)),                       # anchor range overflow
    ]:
        with pytest.raises(ops.YmsError):
            plan.fuse_decode(*args)
    part = torch.zeros(B, H // 2, W // 2, 64, device=DEV)
    with pytest.raises(ops.YmsError):
        ops.ConvPlan(x, w1, bias, yf, ksize=1, act=False).add_upsampled(part)                                # f32 output
    with pytest.raises(ops.YmsError):
        ops.ConvPlan(x, torch.zeros(9, 64, 64, device=DEV, dtype=torch.bfloat16), bias, yb, ksize=3).add_upsampled(part)
    with pytest.raises(ops.YmsError):
        ops.ConvPlan(x, w1, bias, yb, ksize=1).add_upsampled(part.to(torch.bfloat16))


@pytest.mark.parametrize("B,H,W,cout", [(2, 64, 96, 64), (1, 40, 24, 48), (1, 320, 320, 64), (2, 36, 52, 128)])
def test_conv_s2_pair_line_kernel(ops, B, H, W, cout):
    """variant 4: 3x3/s2 with 32 dense input channels on pair-packed weights vs plain PyTorch fp32 conv (and vs the
    generic kernel).  Odd tile counts (H/2 not a multiple of the tile height, W/2 not a multiple of 8) included."""
    g = torch.Generator().manual_seed(B * 1000 + H + W + cout)
    x = torch.randn(B, H, W, 32, generator=g).to(DEV).to(torch.bfloat16)
    wt = (torch.randn(cout, 32, 3, 3, generator=g) / (32 * 9) ** 0.5).to(DEV).to(torch.bfloat16)
    bias = (torch.randn(cout, generator=g) * 0.5).to(DEV)
    wpk = wt.permute(2, 3, 0, 1).reshape(9, cout, 32).contiguous()
    tiles = []
    for ky in range(3):
        tiles.append(torch.cat([wpk[ky * 3 + 1], wpk[ky * 3 + 2]], -1))
        tiles.append(torch.cat([torch.zeros_like(wpk[ky * 3]), wpk[ky * 3]], -1))
    wpair = torch.stack(tiles, 0).contiguous()
    y4 = torch.full((B, H // 2, W // 2, cout), 7.0, device=DEV, dtype=torch.bfloat16)
    y0 = torch.empty_like(y4)
    ops.ConvPlan(x, wpair, bias, y4, ksize=3, stride=2, act=True, variant=4).run()
    ops.ConvPlan(x, wpk, bias, y0, ksize=3, stride=2, act=True).run()
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert rel_l2(y4, ref) < 4e-3
    assert rel_l2(y4, y0.float()) < 3e-3


def test_conv_rejects_cpu_and_bad_shapes(ops):
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16)
    with pytest.raises(ops.YmsError):
        ops.ConvPlan(x, torch.zeros(1, 64, 64, dtype=torch.bfloat16), torch.zeros(64), x.clone())
    xg = torch.zeros(1, 8, 8, 60, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(ops.YmsError):   # channels % 8
        ops.ConvPlan(xg, torch.zeros(1, 64, 60, dtype=torch.bfloat16, device=DEV), torch.zeros(64, device=DEV),
                     torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device=DEV))


@pytest.mark.parametrize("cout", [16, 32, 48, 80])
def test_stem_conv(ops, cout):
    g = torch.Generator().manual_seed(cout)
    x = torch.randn(2, 3, 64, 96, generator=g).to(DEV)
    w = (torch.randn(cout, 3, 3, 3, generator=g) * 0.3).to(DEV)
    b = (torch.randn(cout, generator=g) * 0.2).to(DEV)
    y = torch.empty(2, 32, 48, cout, device=DEV, dtype=torch.bfloat16)
    ops.stem_conv(x, w, b, y)
    # tensor-core stem: image and weights are rounded to bf16 (fp32 accumulate)
    bf = lambda t: t.to(torch.bfloat16).float()
    ref = F.silu(F.conv2d(bf(x), bf(w), b, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < 4e-3
    ref32 = F.silu(F.conv2d(x, w, b, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert rel_l2(y, ref32) < 1e-2


@pytest.mark.parametrize("B,H,W,cout", [(2, 64, 256, 32), (1, 96, 320, 48), (2, 32, 640, 32), (1, 64, 260, 16),
                                          (1, 6, 320, 16)])      # nine half tiles: the last tile has one half past the end of the batch
def test_stem_tma_variant_fp32_and_u8(ops, B, H, W, cout, monkeypatch):
    """Image widths >= 256 take the TMA-fed stem (raw rows through a TMA ring, tiles of two 64-pixel half rows, partial last half when
    W/2 is not a multiple of 64): vs plain PyTorch, fp32 and uint8 inputs, and vs the gather kernel (library option stem_gather)."""
    g = torch.Generator().manual_seed(W + cout)
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1); std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    x = (((img.permute(0, 3, 1, 2).float() / 255.0) - mean) / std).contiguous()
    w = (torch.randn(cout, 3, 3, 3, generator=g) * 0.3).to(DEV); b = (torch.randn(cout, generator=g) * 0.2).to(DEV)
    ref = F.silu(F.conv2d(x.to(DEV), w, b, stride=2, padding=1)).permute(0, 2, 3, 1)
    yf = torch.empty(B, H // 2, W // 2, cout, device=DEV, dtype=torch.bfloat16); yu = torch.empty_like(yf)
    ops.stem_conv(x.to(DEV), w, b, yf)
    ops.stem_conv_u8(img.to(DEV), w, b, yu)
    assert rel_l2(yf, ref) < 1e-2 and rel_l2(yu, ref) < 1e-2
    assert rel_l2(yu, yf.float()) < 2e-3
    from yolo_ms_b200 import _lib
    _lib.set_debug_option("stem_gather", 1)             # same arithmetic, different data path: bit-identical outputs
    try:
        yg = torch.empty_like(yf); ygu = torch.empty_like(yf)
        ops.stem_conv(x.to(DEV), w, b, yg)
        ops.stem_conv_u8(img.to(DEV), w, b, ygu)
    finally:
        _lib.set_debug_option("stem_gather", 0)
    assert torch.equal(yg, yf) and torch.equal(ygu, yu)


def test_stem_tma_many_back_to_back_launches(ops):
    """Regression for the ring-slot ownership bug of the first TMA-fed stem (3 converter groups on 8-deep rings: a group could
    pass the parity wait of a slot two uses ahead): 60 back-to-back launches on a 1280-wide fp32 batch must complete and agree."""
    g = torch.Generator().manual_seed(9)
    x = torch.randn(6, 3, 1280, 1280, generator=g).to(DEV)
    w = (torch.randn(32, 3, 3, 3, generator=g) * 0.3).to(DEV); b = (torch.randn(32, generator=g) * 0.2).to(DEV)
    y0 = torch.empty(6, 640, 640, 32, device=DEV, dtype=torch.bfloat16); y = torch.empty_like(y0)
    ops.stem_conv(x, w, b, y0)
    for _ in range(60):
        ops.stem_conv(x, w, b, y)
    torch.cuda.synchronize()
    assert torch.equal(y, y0)


def test_stem_conv_u8_fuses_totensor_normalize(ops):
    """uint8 HWC image -> (v/255 - mean)/std -> stem, vs the fp32 stem on the reference's own
    ToTensor + Normalize arithmetic (yolov8/tools/test.py:114-119)."""
    g = torch.Generator().manual_seed(5)
    img = torch.randint(0, 256, (2, 64, 96, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1); std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    x = ((img.permute(0, 3, 1, 2).float() / 255.0) - mean) / std                 # ToTensor + Normalize
    w = (torch.randn(32, 3, 3, 3, generator=g) * 0.3).to(DEV); b = (torch.randn(32, generator=g) * 0.2).to(DEV)
    y8 = torch.empty(2, 32, 48, 32, device=DEV, dtype=torch.bfloat16); y32 = torch.empty_like(y8)
    ops.stem_conv_u8(img.to(DEV), w, b, y8)
    ops.stem_conv(x.to(DEV).contiguous(), w, b, y32)
    assert rel_l2(y8, y32.float()) < 2e-3
    ref = F.silu(F.conv2d(x.to(DEV), w, b, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert rel_l2(y8, ref) < 1e-2


def test_sppf_pool_is_exact(ops):
    g = torch.Generator().manual_seed(1)
    for (h, w, c) in ((20, 24, 64), (5, 3, 8), (40, 40, 32), (50, 50, 16)):
        buf = torch.zeros(2, h, w, 4 * c, device=DEV, dtype=torch.bfloat16)
        buf[..., :c] = torch.randn(2, h, w, c, generator=g).to(DEV).to(torch.bfloat16)
        ops.sppf_pool(buf, c)
        x0 = buf[..., :c].float().permute(0, 3, 1, 2)
        x1 = F.max_pool2d(x0, 5, 1, 2); x2 = F.max_pool2d(x1, 5, 1, 2); x3 = F.max_pool2d(x2, 5, 1, 2)
        ref = torch.cat([x0, x1, x2, x3], 1).permute(0, 2, 3, 1)
        assert bool((buf.float() == ref).all())


def test_upsample_into_slice_is_exact(ops):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 10, 12, 64, generator=g).to(DEV).to(torch.bfloat16)
    ybuf = torch.zeros(2, 20, 24, 96, device=DEV, dtype=torch.bfloat16)
    ops.upsample2x(x, ybuf[..., :64])
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert bool((ybuf[..., :64].float() == ref).all()) and bool((ybuf[..., 64:] == 0).all())


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_dwconv(ops, k):
    g = torch.Generator().manual_seed(k)
    for (h, w_, c) in ((20, 20, 64), (37, 45, 24)):
        x = torch.randn(2, h, w_, c, generator=g).to(DEV).to(torch.bfloat16)
        wt = (torch.randn(c, 1, k, k, generator=g) / k).to(DEV)
        b = (torch.randn(c, generator=g) * 0.2).to(DEV)
        y = torch.empty(2, h, w_, c, device=DEV, dtype=torch.bfloat16)
        ops.dwconv(x, wt.reshape(c, k * k).t().contiguous(), b, y, k)
        ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wt, b, padding=k // 2, groups=c)).permute(0, 2, 3, 1)
        assert rel_l2(y, ref) < 4e-3


# ----------------------------------------------------------------------------------------------
# decode / candidate selection vs the oracle
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("sizes,nc", [(((20, 24), (10, 12), (5, 6)), 80), (((15, 15), (7, 9), (3, 5)), 3), (((2, 2), (1, 1), (1, 1)), 80)])
def test_head_decode_matches_oracle(ops, dt, sizes, nc):
    """fp32 decode; tolerance 5e-3 px on boxes, 5e-6 on scores (ex2/rcp approximations).  The emitted
    candidates must be bit-identical to the reference selection applied to OUR prediction."""
    from oracle import postprocess as P
    from oracle import yolov8_oracle as O
    g = torch.Generator().manual_seed(3)
    B = 3
    raw = [(torch.randn(B, 64 + nc, h, w, generator=g) * 2).to(dt).float() for h, w in sizes]
    want = O.decode(raw, (8.0, 16.0, 32.0))
    rawg = [r.permute(0, 2, 3, 1).contiguous().to(DEV).to(dt) for r in raw]
    pred, (cb, cs, cl) = ops.head_decode(rawg, (8.0, 16.0, 32.0), nc, with_candidates=True)
    assert pred.shape == want.shape
    p = pred.cpu()
    assert float((p[..., :4] - want[..., :4]).abs().max()) < 5e-3
    assert float((p[..., 4:] - want[..., 4:]).abs().max()) < 5e-6
    sb, ss, sl = ops.select_candidates(pred)
    for i in range(B):
        wb, ws, wl = P.select_candidates(p[i].numpy())
        for b_, s_, l_ in ((cb, cs, cl), (sb, ss, sl)):
            assert np.array_equal(b_[i].cpu().numpy(), wb)
            assert np.array_equal(s_[i].cpu().numpy(), ws)
            assert np.array_equal(l_[i].cpu().numpy().astype(np.int64), wl)
    # zero strides -> zero boxes (reference default head.stride, yolov8_head.py:79)
    z = ops.head_decode(rawg, (0.0, 0.0, 0.0), nc)
    assert float(z[..., :4].abs().max()) == 0.0


# ----------------------------------------------------------------------------------------------
# NMS: bit-exact keep lists
# ----------------------------------------------------------------------------------------------
def _nms(ops, boxes, scores, labels, conf, iou, nc, nv=None):
    keep, cnt = ops.nms_batched(torch.from_numpy(boxes).to(DEV), torch.from_numpy(scores).to(DEV),
                                torch.from_numpy(labels).to(DEV), conf, iou, nc,
                                None if nv is None else torch.from_numpy(nv).to(DEV))
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    for i in range(keep.shape[0]):
        assert (keep[i, cnt[i]:] == -1).all()
    return [keep[i, :cnt[i]] for i in range(keep.shape[0])]


@pytest.mark.parametrize("name", ["uniform", "clustered", "ties", "degenerate", "exact_thr"])
def test_nms_matches_torchvision_goldens(ops, name):
    g = np.load(os.path.join(GOLDEN, "nms_cases.npz"))
    boxes, scores = g[f"{name}_boxes"], g[f"{name}_scores"]
    labels = np.zeros((1, boxes.shape[0]), np.int32)
    for thr in (0.45, 0.5, 1.0 / 3.0):
        got = _nms(ops, boxes[None], scores[None], labels, -1.0, thr, 1)[0]
        assert np.array_equal(got, g[f"{name}_keep_{thr:.4f}"]), (name, thr)


def test_postprocess_matches_reference_golden(ops):
    from yolo_ms_b200 import postprocess
    g = np.load(os.path.join(GOLDEN, "post_n.npz"))
    pred = torch.from_numpy(g["pred"]).to(DEV)
    for tag in ("a", "b"):
        conf, iou = g[f"thr_{tag}"]
        res = postprocess(pred, float(conf), float(iou))
        for i, (bx, sc, lb) in enumerate(res):
            assert np.array_equal(bx.cpu().numpy(), g[f"boxes_{tag}{i}"])
            assert np.array_equal(sc.cpu().numpy(), g[f"scores_{tag}{i}"])
            assert np.array_equal(lb.cpu().numpy(), g[f"labels_{tag}{i}"])
            assert lb.dtype == torch.int64


@pytest.mark.parametrize("B,N,nc,mode", [(4, 1000, 80, "uni"), (3, 8400, 80, "clu"), (2, 30000, 80, "ties"),
                                          (2, 3000, 1, "clu"), (2, 20000, 3, "clu"), (1, 1, 80, "uni"),
                                          (2, 33, 5, "uni"), (2, 16385, 80, "uni"),
                                          (2, 33600, 80, "skew"), (3, 8400, 80, "skewbig"), (40, 2000, 80, "skew"), (2, 6000, 300, "uni"),
                                          (2, 33600, 80, "one9k"), (2, 33600, 80, "one13k"), (2, 33600, 80, "one20k"), (1, 40000, 2, "one30k")])
def test_nms_random_ragged_matches_c_oracle(ops, B, N, nc, mode):
    from oracle import postprocess as P
    rng = np.random.default_rng(N + nc)
    if mode == "clu" and N >= 100:
        ctr = rng.uniform(0, 600, (B, N // 100 + 1, 2))
        xy = np.repeat(ctr, 100, 1)[:, :N] + rng.normal(0, 6, (B, N, 2))
    else:
        xy = rng.uniform(0, 600, (B, N, 2))
    wh = rng.uniform(4, 64, (B, N, 2))
    boxes = np.concatenate([xy, xy + wh], -1).astype(np.float32)
    scores = rng.uniform(0, 1, (B, N)).astype(np.float32)
    if mode == "ties":
        scores = (np.round(scores * 256) / 256).astype(np.float32)
    labels = rng.integers(0, nc, (B, N)).astype(np.int32)
    if mode.startswith("skew"):       # a few dominant classes (what a random-weight head produces): exercises the cost- and
        p = 1.0 / np.arange(1, nc + 1) ** 1.3   # count-balanced class ranges and classes of thousands of boxes
        labels = rng.choice(nc, size=(B, N), p=p / p.sum()).astype(np.int32)
    if mode.startswith("one"):        # ONE dominant class of 9k / 13k / 20k / 30k boxes next to uniform ones (MS-Block legs of bench.py at
        big = int(mode[3:-1]) * 1000  # 1280 x 1280): boxes beyond shared memory (> 8192), keys beyond shared memory (> 16384)
        labels[:, :big] = 0
        wh = rng.uniform(30, 120, (B, N, 2))
        boxes = np.concatenate([xy, xy + wh], -1).astype(np.float32)
    if mode == "skewbig":             # large, heavily overlapping boxes (15 x stride wide), like the bench workload
        wh = rng.uniform(100, 480, (B, N, 2))
        boxes = np.concatenate([xy - wh / 2, xy + wh / 2], -1).astype(np.float32)
    nv = rng.integers(N // 2, N + 1, (B,)).astype(np.int32)
    got = _nms(ops, boxes, scores, labels, 0.25, 0.45, nc, nv)
    for i in range(B):
        want = P.class_nms_c(boxes[i, :nv[i]], scores[i, :nv[i]], labels[i, :nv[i]], 0.25, 0.45)
        assert np.array_equal(got[i], want)


def test_nms_edge_cases(ops):
    boxes = np.array([[[0, 0, 10, 10], [0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 20, 20]]], np.float32)
    scores = np.array([[0.25, 0.9, 0.8, 0.7]], np.float32)
    labels = np.array([[0, 1, 1, 1]], np.int32)
    # score == conf dropped (strict >); identical boxes of one class suppress; zero-area box survives (NaN IoU)
    assert _nms(ops, boxes, scores, labels, 0.25, 0.45, 2)[0].tolist() == [1, 3]
    # nothing above conf
    assert _nms(ops, boxes, scores, labels, 0.95, 0.45, 2)[0].tolist() == []
    # n_valid = 0 and out-of-range labels are ignored
    assert _nms(ops, boxes, scores, labels, 0.1, 0.45, 2, np.array([0], np.int32))[0].tolist() == []
    assert _nms(ops, boxes, scores, np.array([[0, 5, -1, 1]], np.int32), 0.1, 0.45, 2)[0].tolist() == [0, 3]
    # iou threshold 1.0: nothing is ever suppressed (IoU <= 1)
    assert _nms(ops, boxes, scores, labels, 0.1, 1.0, 2)[0].tolist() == [0, 1, 2, 3]
    # empty batch / empty N
    k, c = ops.nms_batched(torch.zeros(2, 0, 4, device=DEV), torch.zeros(2, 0, device=DEV),
                           torch.zeros(2, 0, dtype=torch.int32, device=DEV), 0.25, 0.45, 80)
    assert k.shape == (2, 0) and c.tolist() == [0, 0]


def test_nms_full_size_properties(ops):
    """BASELINE config 4 (30k boxes x batch 64 x 80 classes): size-independent properties --
    idempotence, output ordering, no surviving same-class pair above the threshold -- plus the C
    oracle on two images of the batch."""
    from oracle import postprocess as P
    from gpu_util import iou_xyxy
    rng = np.random.default_rng(0)
    B, N, nc = 64, 30000, 80
    xy = rng.uniform(0, 600, (B, N, 2)); wh = rng.uniform(4, 64, (B, N, 2))
    boxes = np.concatenate([xy, xy + wh], -1).astype(np.float32)
    scores = rng.uniform(0, 1, (B, N)).astype(np.float32)
    labels = rng.integers(0, nc, (B, N)).astype(np.int32)
    got = _nms(ops, boxes, scores, labels, 0.25, 0.45, nc)
    for i in (0, 63):
        assert np.array_equal(got[i], P.class_nms_c(boxes[i], scores[i], labels[i], 0.25, 0.45))
    for i in (5, 31):
        k = got[i]
        key = labels[i][k].astype(np.int64) * 4 - scores[i][k].astype(np.float64)     # label asc, score desc
        assert (np.diff(key) >= 0).all()
        assert (scores[i][k] > 0.25).all()
        # idempotence: NMS over the survivors keeps all of them, in the same order
        again = _nms(ops, boxes[i][k][None], scores[i][k][None], labels[i][k][None], 0.25, 0.45, nc)[0]
        assert np.array_equal(again, np.arange(k.size))
        # sampled pair check inside one class
        sel = k[labels[i][k] == 7]
        a, b = np.meshgrid(np.arange(sel.size), np.arange(sel.size))
        m = a < b
        assert (iou_xyxy(boxes[i][sel][a[m]], boxes[i][sel][b[m]]) <= 0.45 + 1e-6).all()
