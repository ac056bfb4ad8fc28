"""Program builder / executor: turns a module tree into a static list of kernel launches over
pre-allocated NHWC bf16 buffers and replays it (CUDA graph) per call.

Data layout in HBM
------------------
* activations: NHWC bf16, one buffer per tensor, EXCEPT that every tensor that the reference
  concatenates (C2f's [b_n..b_1,x1,x2], SPPF's [x,x1,x2,x3], the neck's [up(x),skip] /
  [conv(x),skip]) is produced directly inside its channel slice of the concat buffer --
  producers get a strided view (pixel stride = concat width), so torch.cat never runs.
* weights: BN folded in fp32, then packed bf16 [tap][c_out][c_in] (K contiguous) = the B operand
  tiles TMA fetches; bias fp32.
* head raw logits: fp32 [B,H,W,64+nc] per scale (box | cls), decode output fp32 [B,A,4+nc].
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch

from . import ops
from ._lib import YmsError

BN_EPS_DEFAULT = 1e-3
AUTOTUNE = os.environ.get("YMS_AUTOTUNE", "1") != "0"     # per-layer kernel-variant selection at program build


def fold_conv_bn(weight: torch.Tensor, bn_w, bn_b, bn_mean, bn_var, eps: float):
    """conv(bias=False)+BN(eval) -> (weight', bias') in fp32 (components.py:69-77 of the reference)."""
    scale = bn_w.float() / torch.sqrt(bn_var.float() + eps)
    return weight.float() * scale.view(-1, 1, 1, 1), bn_b.float() - bn_mean.float() * scale


def pack_weight(w: torch.Tensor) -> torch.Tensor:
    """[c_out, c_in, k, k] fp32 -> bf16 [k*k, c_out, c_in] (tap-major, K contiguous)."""
    co, ci, k, _ = w.shape
    return w.permute(2, 3, 0, 1).reshape(k * k, co, ci).contiguous().to(torch.bfloat16)


# Autotuner decisions, shared by every program of the process: a layer is identified by its geometry and buffer strides, so a
# second program over the same layers (another input dtype, another module instance) reuses the measured choice instead of
# re-timing it -- possibly under interference from concurrent copies or kernels -- and the build is deterministic per process.
_TUNE_CACHE: dict = {}


def save_tune_cache(path: str) -> None:
    """Write the autotuner's decisions of this process (layer geometry -> chosen kernel variant) as text."""
    with open(path, "w") as f:
        f.write(repr(sorted(_TUNE_CACHE.items(), key=repr)))


def load_tune_cache(path: str) -> int:
    """Adopt decisions saved by save_tune_cache (e.g. so that a run under a profiler, whose timings are perturbed, builds the
    SAME programs as the plain run before it).  Returns the number of entries read."""
    import ast
    with open(path) as f:
        items = ast.literal_eval(f.read())
    _TUNE_CACHE.update(dict(items))
    return len(items)


def _tune_key(kind, x, y, ksize, stride, act, residual):
    return (kind, str(x.device), tuple(x.shape), x.stride(-2), tuple(y.shape), y.stride(-2), ksize, stride, bool(act),
            None if residual is None else residual.stride(-2))


class Program:
    """A recorded sequence of launches over static buffers."""

    def __init__(self, device: torch.device):
        self.device = device
        self.steps: List[Callable[[], None]] = []
        self.plans: List[ops.ConvPlan] = []
        self._keep: list = []
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.eager_prefix = 0          # leading steps that read caller-owned memory: never captured
        self.schedule: list = []       # ("op", fn, branch) | ("fork", n, 0) | ("join", 0, 0)
        self._branch = 0
        self._n_side = 0
        self._side: List[torch.cuda.Stream] = []
        self.launches = 0
        self.names: List[str] = []     # one label per step (profiling / per-layer tables)
        self.costs: dict = {}          # step index -> (flops, bytes) of non-conv steps
        self.tuned: list = []          # (layer, chosen variant) of the autotuned 3x3 layers
        self.flops = 0.0
        self.bytes = 0.0
        self.raw_tail: List[ops.ConvPlan] = []   # decode-fused programs: the plans that store the raw head logits instead
        self.decoded: Optional[dict] = None      # decode-fused programs: static pred / candidate buffers + the stride tensor
        self.post: dict = {}                     # max_det -> ops.PostBuffers (static NMS / gather outputs of YOLOv8.detect)
        self.full_graphs: dict = {}              # (input address, conf, iou, max_det) -> (graph of the WHOLE step, outputs)
        self.full_keep: list = []
        self.seen: dict = {}

    # ---- buffers -------------------------------------------------------------------------
    def buf(self, b: int, h: int, w: int, c: int, dtype=torch.bfloat16) -> torch.Tensor:
        t = torch.empty((b, h, w, c), dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def hold(self, *ts):
        self._keep.extend(ts)

    # ---- ops -------------------------------------------------------------------------------
    def conv(self, weight_packed, bias, x, y, ksize, stride=1, act=True, residual=None, x2=None, decode=None, scheduled=True,
             up_add=None):
        """decode: kwargs of ConvPlan.fuse_decode (the head's final convs of the fused forward+decode program).
        scheduled=False: the plan is built and kept but is not part of the program (returned to the caller)."""
        plan = ops.ConvPlan(x, weight_packed, bias, y, ksize=ksize, stride=stride, act=act, residual=residual, x2=x2)
        if decode is not None:
            plan.fuse_decode(**decode)
        if up_add is not None:           # fp32 partial sums at half resolution (ConvPlan.add_upsampled)
            plan.add_upsampled(up_add)
        if not scheduled:
            self.hold(weight_packed, bias)
            return plan
        if (AUTOTUNE and ksize == 3 and stride == 2 and x2 is None and residual is None and y.dtype == torch.bfloat16
                and x.shape[-1] == 32 and x.stride(-2) == 32 and y.shape[-1] <= 256):
            plan = self._autotune_s2pair(plan, weight_packed, dict(x=x, bias=bias, y=y, act=act))
        if AUTOTUNE and ksize == 3 and stride == 1 and x2 is None and y.dtype == torch.bfloat16:
            plan = self._autotune(plan, dict(x=x, weight=weight_packed, bias=bias, y=y, ksize=ksize, stride=stride, act=act,
                                             residual=residual, x2=x2))
        elif AUTOTUNE and y.dtype == torch.bfloat16 and up_add is None and plan.variant == 0:
            # 1x1 / 3x3-s2 / two-source layers on the generic kernel: single CTA or CTA pair (variant 5)
            plan = self._autotune(plan, dict(x=x, weight=weight_packed, bias=bias, y=y, ksize=ksize, stride=stride, act=act,
                                             residual=residual, x2=x2), variants=(5,), tag=f"pair{0 if x2 is None else x2.shape[-1]}")
        self.plans.append(plan)
        self.hold(weight_packed, bias)
        self._push(plan.run, plan.desc)
        self.flops += plan.flops
        self.bytes += plan.bytes
        return y

    @staticmethod
    def _time_plan(plan, reps=5):
        for _ in range(2):
            plan.run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            plan.run()
        b.record()
        b.synchronize()
        return a.elapsed_time(b)

    def _autotune_s2pair(self, default_plan, weight_packed, kw):
        """3x3/s2 with 32 dense input channels (backbone.conv1 of the s model): the pair-line kernel (variant 4) reads
        each input pixel pair once instead of nine shifted tiles.  Its weights are pair-packed [6][c_out][64]:
        tile 2*ky = [w(ky,1) | w(ky,2)], tile 2*ky+1 = [0 | w(ky,0)]."""
        key = _tune_key("s2pair", kw["x"], kw["y"], 3, 2, kw["act"], None)
        if _TUNE_CACHE.get(key) == 0:
            return default_plan
        w = weight_packed                                         # [9, c_out, 32], tap = ky*3 + kx
        tiles = []
        for ky in range(3):
            tiles.append(torch.cat([w[ky * 3 + 1], w[ky * 3 + 2]], dim=-1))
            tiles.append(torch.cat([torch.zeros_like(w[ky * 3]), w[ky * 3]], dim=-1))
        wp = torch.stack(tiles, 0).contiguous()
        try:
            cand = ops.ConvPlan(kw["x"], wp, kw["bias"], kw["y"], ksize=3, stride=2, act=kw["act"], variant=4)
        except YmsError:
            return default_plan
        if key not in _TUNE_CACHE:
            _TUNE_CACHE[key] = 4 if self._time_plan(cand) < 0.97 * self._time_plan(default_plan) else 0
        if _TUNE_CACHE[key] == 4:
            self.hold(wp)
            cand.desc = default_plan.desc + " [v4]"
            self.tuned.append((default_plan.desc, 4))
            return cand
        return default_plan

    def _autotune(self, default_plan, kw, variants=(1, 2, 3, 5, 6, 7), tag="3x3"):
        """Measure, don't guess: the 3x3/s1 layers have three tcgen05 implementations whose winner depends on the map
        size (tile quantisation on 20x20 / 40x40 maps), c_out (resident vs streamed weights) and the tile count per CTA.
        Each candidate runs on the layer's real buffers at program-build time; the fastest one is kept."""
        def timed(plan):
            for _ in range(2):
                plan.run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                plan.run()
            b.record()
            b.synchronize()
            return a.elapsed_time(b)
        key = _tune_key(tag, kw["x"], kw["y"], kw["ksize"], kw["stride"], kw["act"], kw["residual"])
        known = _TUNE_CACHE.get(key)
        if known is not None:                                     # decided earlier in this process: rebuild that variant, no timing
            best = default_plan
            if known != default_plan.variant:
                try:
                    best = ops.ConvPlan(kw["x"], kw["weight"], kw["bias"], kw["y"], ksize=kw["ksize"], stride=kw["stride"], act=kw["act"],
                                        residual=kw["residual"], x2=kw["x2"], variant=known)
                except YmsError:
                    best = default_plan
            self.tuned.append((default_plan.desc, best.variant))
            best.desc = default_plan.desc + (f" [v{best.variant}]" if best.variant else "")
            return best
        best, best_t = default_plan, timed(default_plan)
        for variant in variants:
            try:
                cand = ops.ConvPlan(kw["x"], kw["weight"], kw["bias"], kw["y"], ksize=kw["ksize"], stride=kw["stride"], act=kw["act"],
                                    residual=kw["residual"], x2=kw["x2"], variant=variant)
            except YmsError:
                continue
            t = timed(cand)
            if t < 0.97 * best_t:
                best, best_t = cand, t
        _TUNE_CACHE[key] = best.variant
        self.tuned.append((default_plan.desc, best.variant))
        best.desc = default_plan.desc + (f" [v{best.variant}]" if best.variant else "")
        return best

    def ms_layer(self, plan):
        """A ready-made plan (ops.MsLayerPlan / ops.ConvPlan): one launch, costed by the plan itself."""
        self.plans.append(plan)
        self.hold(*[t for t in plan._keep if t is not None])
        self._push(plan.run, plan.desc)
        self.flops += plan.flops
        self.bytes += plan.bytes

    def pick_fastest(self, key, candidates):
        """candidates: [(tag, [plans])], most fused first.  With the autotuner on, every candidate sequence is timed on the
        layer's real buffers (decision cached per process under `key`); otherwise the first one is taken."""
        if not AUTOTUNE or len(candidates) == 1:
            return candidates[0]
        known = _TUNE_CACHE.get(key)
        if known is not None:
            for c in candidates:
                if c[0] == known:
                    return c
        def timed(plans):
            for _ in range(2):
                for pl in plans:
                    pl.run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                for pl in plans:
                    pl.run()
            b.record()
            b.synchronize()
            return a.elapsed_time(b)
        best, best_t = candidates[0], timed(candidates[0][1])
        for c in candidates[1:]:
            t = timed(c[1])
            if t < 0.97 * best_t:
                best, best_t = c, t
        _TUNE_CACHE[key] = best[0]
        self.tuned.append((key, best[0]))
        return best

    def add(self, fn: Callable[[], None], nbytes: float = 0.0, flops: float = 0.0, name: str = "op"):
        self._push(fn, name)
        self.costs[len(self.steps) - 1] = (flops, nbytes)
        self.bytes += nbytes
        self.flops += flops

    def _push(self, fn, name="op"):
        self.steps.append(fn)
        self.names.append(name)
        self.schedule.append(("op", fn, self._branch))
        self.launches += 1

    # ---- independent branches (e.g. the three head scales): captured as parallel graph branches ----
    def fork(self, n_branches: int):
        self.schedule.append(("fork", n_branches, 0))
        self._n_side = max(self._n_side, n_branches - 1)

    def fork_one(self, k: int):
        """Side branch k (> 0) starts here: its stream waits for everything the main branch has issued so far."""
        self.schedule.append(("fork_one", k, 0))
        self._n_side = max(self._n_side, k)

    def branch(self, i: int):
        self._branch = i

    def join(self):
        self.schedule.append(("join", 0, 0))
        self._branch = 0

    # ---- execution ---------------------------------------------------------------------------
    def run_eager(self, start: int = 0):
        """Issue every launch (from op index `start`) on the current stream; branches > 0 go to side
        streams between fork/join."""
        main = torch.cuda.current_stream(self.device)
        while len(self._side) < self._n_side:
            self._side.append(torch.cuda.Stream(device=self.device))
        op_idx = 0
        used = set()
        for kind, arg, br in self.schedule:
            if kind == "op":
                if op_idx >= start:
                    if br == 0:
                        arg()
                    else:
                        used.add(br)
                        with torch.cuda.stream(self._side[br - 1]):
                            arg()
                op_idx += 1
            elif kind == "fork":
                if op_idx >= start:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    for k in range(arg - 1):
                        self._side[k].wait_event(ev)
            elif kind == "fork_one":
                if op_idx >= start:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    self._side[arg - 1].wait_event(ev)
            elif kind == "join":
                for k in sorted(used):
                    ev = torch.cuda.Event()
                    ev.record(self._side[k - 1])
                    main.wait_event(ev)
                used.clear()

    def capture(self):
        """Warm up once, then capture steps[eager_prefix:] into a CUDA graph."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.run_eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_eager(start=self.eager_prefix)
        self.graph = g

    def run(self):
        if self.graph is not None:
            for s in self.steps[:self.eager_prefix]:
                s()
            self.graph.replay()
        else:
            self.run_eager()


def nchw_f32_to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """API adaptation for stand-alone sub-modules (NOT on the YOLOv8.forward hot path, where the
    stem kernel reads the NCHW fp32 image directly)."""
    if not x.is_cuda:
        raise YmsError("yolo_ms_b200 modules run on CUDA tensors only (no CPU fallback)")
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nhwc_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).float()
