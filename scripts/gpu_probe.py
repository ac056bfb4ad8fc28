"""Kernel bring-up probe for the B200 box: runs every kernel against a reference, case by case,
each group in a subprocess (a trapped kernel kills only that subprocess), and writes a report
to gpurun_out/probe.log.  Not part of the test-suite; use `pytest -m gpu` for the gated tests.

    python scripts/gpu_probe.py            # all groups
    python scripts/gpu_probe.py conv 5     # (child) run conv cases from index 5
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

# name, B, H, W, cin, cout, k, stride, act, res, cin2, out_f32, slices
CONV_CASES = [
    ("1x1_64_64", 2, 40, 40, 64, 64, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_128_256", 2, 40, 40, 128, 256, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_tailM", 1, 20, 20, 64, 64, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_32_32", 2, 16, 16, 32, 32, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_16_16", 2, 16, 16, 16, 16, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_48_96", 2, 16, 16, 48, 96, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_80_80_f32", 2, 20, 20, 80, 80, 1, 1, 0, 0, 0, 1, 0),
    ("1x1_64_64_f32_slice", 2, 20, 20, 64, 64, 1, 1, 0, 0, 0, 1, 1),
    ("1x1_768_512", 2, 20, 20, 768, 512, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_576_576", 1, 20, 20, 576, 576, 1, 1, 1, 0, 0, 0, 0),
    ("1x1_slices", 2, 40, 40, 64, 64, 1, 1, 1, 0, 0, 0, 1),
    ("1x1_two_src", 2, 40, 40, 64, 64, 1, 1, 1, 0, 64, 0, 0),
    ("1x1_two_src_odd", 2, 20, 20, 48, 80, 1, 1, 1, 0, 32, 0, 1),
    ("3x3_64_64", 2, 40, 40, 64, 64, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_64_64_res", 2, 40, 40, 64, 64, 3, 1, 1, 1, 0, 0, 0),
    ("3x3_res_slices", 2, 20, 20, 128, 128, 3, 1, 1, 1, 0, 0, 1),
    ("3x3_128_80_w20", 2, 20, 20, 128, 80, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_32_32_w160", 1, 160, 160, 32, 32, 3, 1, 1, 0, 0, 0, 0),
    ("3x3_80_80_w80", 1, 80, 80, 80, 80, 3, 1, 1, 0, 0, 0, 0),
    ("3x3s2_64_128", 2, 80, 80, 64, 128, 3, 2, 1, 0, 0, 0, 0),
    ("3x3s2_32_64", 1, 160, 160, 32, 64, 3, 2, 1, 0, 0, 0, 0),
    ("3x3s2_slices", 2, 40, 40, 128, 128, 3, 2, 1, 0, 0, 0, 1),
    ("3x3s2_odd_out", 1, 24, 40, 64, 64, 3, 2, 1, 0, 0, 0, 0),
    ("3x3_512_64_big", 4, 80, 80, 128, 144, 3, 1, 1, 0, 0, 0, 0),
]


def child_conv(start: int):
    import torch
    import torch.nn.functional as F
    from yolo_ms_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    for ci in range(start, len(CONV_CASES)):
        name, B, H, W, cin, cout, k, s, act, res, cin2, f32, sl = CONV_CASES[ci]
        print(f"CASE {ci} {name} BEGIN", flush=True)
        g = torch.Generator(device="cpu").manual_seed(100 + ci)
        pad_c = 24 if sl else 0

        def mk(c, h, w):
            full = (torch.randn(B, h, w, c + pad_c, generator=g)).to(dev).to(torch.bfloat16)
            return full, full[..., 8:8 + c] if sl else full

        xf, x = mk(cin, H, W)
        x2f, x2 = mk(cin2, H, W) if cin2 else (None, None)
        ktot = cin + cin2
        wt = (torch.randn(cout, ktot, k, k, generator=g) / (ktot * k * k) ** 0.5).to(dev).to(torch.bfloat16)
        bias = torch.randn(cout, generator=g).to(dev) * 0.5
        Ho, Wo = H // s, W // s
        rf, r = mk(cout, Ho, Wo) if res else (None, None)
        odt = torch.float32 if f32 else torch.bfloat16
        yfull = torch.full((B, Ho, Wo, cout + pad_c), 7.0, device=dev, dtype=odt)
        y = yfull[..., 8:8 + cout] if sl else yfull
        wpk = wt.permute(2, 3, 0, 1).reshape(k * k, cout, ktot).contiguous()
        try:
            plan = ops.ConvPlan(x, wpk, bias, y, ksize=k, stride=s, act=bool(act), residual=r, x2=x2)
            plan.run()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"CASE {ci} {name} ERROR {type(e).__name__}: {e}", flush=True)
            if "CUDA" in str(e) or "cuda" in str(e):
                raise
            continue
        xin = x.float() if x2 is None else torch.cat([x.float(), x2.float()], -1)
        ref = F.conv2d(xin.permute(0, 3, 1, 2), wt.float(), bias, stride=s, padding=k // 2)
        if act:
            ref = F.silu(ref)
        if res:
            ref = ref + r.float().permute(0, 3, 1, 2)
        ref = ref.permute(0, 2, 3, 1)
        got = y.float()
        err = (got - ref).abs()
        scale = float(ref.abs().max())
        rel = float(err.max()) / scale
        rl2 = float((got - ref).norm() / ref.norm())
        untouched = True
        if sl:
            untouched = bool((yfull[..., :8] == 7).all() and (yfull[..., 8 + cout:] == 7).all())
        ok = rel < 2e-2 and rl2 < 1e-2 and untouched
        extra = ""
        if not ok:
            e2 = err.amax(dim=(0, 3))          # [Ho, Wo]
            bad = (e2 > 2e-2 * scale).nonzero()
            ec = err.amax(dim=(0, 1, 2))
            badc = (ec > 2e-2 * scale).nonzero().flatten()
            extra = f" badpix={bad[:12].tolist()} n_badpix={bad.shape[0]} badch={badc[:16].tolist()} n_badch={badc.numel()} got0={got.flatten()[:4].tolist()} ref0={ref.flatten()[:4].tolist()}"
        print(f"CASE {ci} {name} {'OK' if ok else 'FAIL'} maxrel={rel:.3e} relL2={rl2:.3e} untouched={untouched}"
              f" flops={plan.flops:.3g}{extra}", flush=True)
    print("CONV DONE", flush=True)


def child_post():
    import numpy as np
    import torch
    from oracle import postprocess as P
    from oracle import yolov8_oracle as O
    from yolo_ms_b200 import ops
    dev = "cuda"
    # ---- decode ----
    for dt in (torch.float32, torch.bfloat16):
        g = torch.Generator().manual_seed(3)
        B, nc = 3, 80
        sizes = [(20, 24), (10, 12), (5, 6)]
        raw = [(torch.randn(B, 64 + nc, h, w, generator=g) * 2).to(dt).float() for h, w in sizes]
        want = O.decode(raw, (8.0, 16.0, 32.0))
        rawg = [r.permute(0, 2, 3, 1).contiguous().to(dev).to(dt) for r in raw]
        pred, (cb, cs, cl) = ops.head_decode(rawg, (8.0, 16.0, 32.0), nc, with_candidates=True)
        torch.cuda.synchronize()
        pred = pred.cpu()
        eb = float((pred[..., :4] - want[..., :4]).abs().max())
        es = float((pred[..., 4:] - want[..., 4:]).abs().max())
        wb, ws, wl = zip(*[P.select_candidates(pred[i].numpy()) for i in range(B)])
        okc = all(np.array_equal(cb[i].cpu().numpy(), wb[i]) and np.array_equal(cs[i].cpu().numpy(), ws[i])
                  and np.array_equal(cl[i].cpu().numpy().astype(np.int64), wl[i]) for i in range(B))
        print(f"DECODE {dt} box_err={eb:.3e} score_err={es:.3e} cand_exact={okc}", flush=True)
        sb, ss, sl = ops.select_candidates(pred.to(dev))
        oks = all(np.array_equal(sb[i].cpu().numpy(), wb[i]) and np.array_equal(ss[i].cpu().numpy(), ws[i])
                  and np.array_equal(sl[i].cpu().numpy().astype(np.int64), wl[i]) for i in range(B))
        print(f"SELECT exact={oks}", flush=True)
    # ---- NMS goldens ----
    g = np.load(os.path.join(ROOT, "tests", "golden", "nms_cases.npz"))
    for name in ("uniform", "clustered", "ties", "degenerate", "exact_thr"):
        boxes, scores = g[f"{name}_boxes"], g[f"{name}_scores"]
        for thr in (0.45, 0.5, 1.0 / 3.0):
            want = g[f"{name}_keep_{thr:.4f}"]
            keep, cnt = ops.nms_batched(torch.from_numpy(boxes)[None].to(dev), torch.from_numpy(scores)[None].to(dev),
                                        torch.zeros(1, boxes.shape[0], dtype=torch.int32, device=dev), -1.0, thr, 1)
            torch.cuda.synchronize()
            got = keep[0, :int(cnt[0])].cpu().numpy()
            print(f"NMS golden {name} thr={thr:.4f} exact={np.array_equal(got, want)} kept={got.size}/{want.size}", flush=True)
    gp = np.load(os.path.join(ROOT, "tests", "golden", "post_n.npz"))
    pred = torch.from_numpy(gp["pred"]).to(dev)
    b, s, l = ops.select_candidates(pred)
    for tag in ("a", "b"):
        conf, iou = gp[f"thr_{tag}"]
        keep, cnt = ops.nms_batched(b, s, l, float(conf), float(iou), 80)
        for i in range(pred.shape[0]):
            got = keep[i, :int(cnt[i])].cpu().numpy()
            print(f"POST golden {tag}{i} exact={np.array_equal(got, gp[f'keep_{tag}{i}'])}", flush=True)
    # ---- NMS random multi-class, ragged, big ----
    rng = np.random.default_rng(5)
    for (B, N, nc, mode) in ((4, 1000, 80, "uni"), (3, 8400, 80, "clu"), (2, 30000, 80, "uni"), (2, 30000, 80, "ties"),
                             (2, 5000, 1, "clu"), (2, 40000, 3, "clu"), (1, 1, 80, "uni"), (2, 33, 5, "uni")):
        if mode == "clu":
            ctr = rng.uniform(0, 600, (B, max(N // 100, 1), 2))
            xy = np.repeat(ctr, 100, 1)[:, :N] + rng.normal(0, 6, (B, N, 2)) if N >= 100 else rng.uniform(0, 600, (B, N, 2))
        else:
            xy = rng.uniform(0, 600, (B, N, 2))
        wh = rng.uniform(4, 64, (B, N, 2))
        boxes = np.concatenate([xy, xy + wh], -1).astype(np.float32)
        scores = rng.uniform(0, 1, (B, N)).astype(np.float32)
        if mode == "ties":
            scores = (np.round(scores * 256) / 256).astype(np.float32)
        labels = rng.integers(0, nc, (B, N)).astype(np.int32)
        nv = rng.integers(N // 2, N + 1, (B,)).astype(np.int32)
        t0 = time.time()
        keep, cnt = ops.nms_batched(torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev),
                                    torch.from_numpy(labels).to(dev), 0.25, 0.45, nc, torch.from_numpy(nv).to(dev))
        torch.cuda.synchronize()
        dt_ms = (time.time() - t0) * 1e3
        ok = True
        for i in range(B):
            want = P.class_nms_c(boxes[i, :nv[i]], scores[i, :nv[i]], labels[i, :nv[i]], 0.25, 0.45)
            got = keep[i, :int(cnt[i])].cpu().numpy()
            ok &= bool(np.array_equal(got, want))
            ok &= bool((keep[i, int(cnt[i]):] == -1).all())
        print(f"NMS random B={B} N={N} nc={nc} {mode} exact={ok} kept={cnt.tolist()} ms={dt_ms:.2f}", flush=True)
    print("POST DONE", flush=True)


def child_glue():
    import torch
    import torch.nn.functional as F
    from yolo_ms_b200 import ops
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    # stem
    for cout in (16, 32, 48):
        x = torch.randn(2, 3, 64, 96, generator=g).to(dev)
        w = (torch.randn(cout, 3, 3, 3, generator=g) * 0.3).to(dev)
        b = torch.randn(cout, generator=g).to(dev) * 0.2
        y = torch.empty(2, 32, 48, cout, device=dev, dtype=torch.bfloat16)
        ops.stem_conv(x, w, b, y)
        ref = F.silu(F.conv2d(x, w, b, stride=2, padding=1)).permute(0, 2, 3, 1)
        print(f"STEM cout={cout} relL2={float((y.float() - ref).norm() / ref.norm()):.3e}", flush=True)
    # sppf
    c = 64
    buf = torch.zeros(2, 20, 24, 4 * c, device=dev, dtype=torch.bfloat16)
    buf[..., :c] = torch.randn(2, 20, 24, c, generator=g).to(dev).to(torch.bfloat16)
    ops.sppf_pool(buf, c)
    x0 = buf[..., :c].float().permute(0, 3, 1, 2)
    x1 = F.max_pool2d(x0, 5, 1, 2); x2 = F.max_pool2d(x1, 5, 1, 2); x3 = F.max_pool2d(x2, 5, 1, 2)
    ref = torch.cat([x0, x1, x2, x3], 1).permute(0, 2, 3, 1)
    print(f"SPPF exact={bool((buf.float() == ref).all())}", flush=True)
    # upsample into slice
    x = torch.randn(2, 10, 12, 64, generator=g).to(dev).to(torch.bfloat16)
    ybuf = torch.zeros(2, 20, 24, 96, device=dev, dtype=torch.bfloat16)
    ops.upsample2x(x, ybuf[..., :64])
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    print(f"UPSAMPLE exact={bool((ybuf[..., :64].float() == ref).all())} rest_zero={bool((ybuf[..., 64:] == 0).all())}", flush=True)
    # depthwise
    for k in (3, 5, 7, 9):
        for (h, w_, c) in ((20, 20, 64), (37, 45, 24)):
            x = torch.randn(2, h, w_, c, generator=g).to(dev).to(torch.bfloat16)
            wt = (torch.randn(c, 1, k, k, generator=g) / k).to(dev)
            b = torch.randn(c, generator=g).to(dev) * 0.2
            y = torch.empty(2, h, w_, c, device=dev, dtype=torch.bfloat16)
            ops.dwconv(x, wt.reshape(c, k * k).t().contiguous(), b, y, k)
            ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wt, b, padding=k // 2, groups=c)).permute(0, 2, 3, 1)
            print(f"DWCONV k={k} {h}x{w_}x{c} relL2={float((y.float() - ref).norm() / ref.norm()):.3e}", flush=True)
    print("GLUE DONE", flush=True)


def run_group(name, log, extra=()):
    cmd = [sys.executable, os.path.abspath(__file__), name, *map(str, extra)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    log.write(f"\n===== {' '.join(cmd[2:])} rc={p.returncode}\n{p.stdout}\n")
    log.flush()
    return p


def main():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "probe.log"), "w") as log:
        for grp in ("post", "glue"):
            try:
                run_group(grp, log)
            except subprocess.TimeoutExpired:
                log.write(f"{grp}: TIMEOUT\n")
        start = 0
        while start < len(CONV_CASES):
            try:
                p = run_group("conv", log, (start,))
            except subprocess.TimeoutExpired:
                log.write(f"conv from {start}: TIMEOUT\n")
                break
            if "CONV DONE" in p.stdout:
                break
            last = [ln for ln in p.stdout.splitlines() if ln.startswith("CASE") and "BEGIN" in ln]
            nxt = int(last[-1].split()[1]) + 1 if last else start + 1
            log.write(f"conv subprocess died in case {nxt - 1}; resuming at {nxt}\n")
            start = nxt
    print(open(os.path.join(OUT, "probe.log")).read()[-6000:])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        {"conv": lambda: child_conv(int(sys.argv[2]) if len(sys.argv) > 2 else 0),
         "post": child_post, "glue": child_glue}[sys.argv[1]]()
    else:
        main()
