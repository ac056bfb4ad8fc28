"""CPU, gloo, world_size 2: the host-side multi-GPU logic (batch sharding, detection all-gather,
max-over-ranks timing)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolo_ms_b200.dist import all_gather_detections, max_over_ranks, shard_range
    lo, hi = shard_range(7, rank, world)
    dets = torch.full((4, 5, 6), float(rank)); dets[:, :, 5] = torch.arange(5.0)
    counts = torch.tensor([rank, rank + 1, rank + 2, rank + 3], dtype=torch.int32)
    d, c = all_gather_detections(dets, counts)
    t = max_over_ranks(1.0 + rank, "cpu")
    q.put((rank, lo, hi, tuple(d.shape), d[:, 0, 0].tolist(), c.tolist(), t))
    dist.destroy_process_group()


def test_shard_gather_and_max_over_ranks():
    from yolo_ms_b200.dist import shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(256, 7, 8) == (224, 256)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 4, 4, 7)
    for r in res:
        assert r[3] == (8, 5, 6)
        assert r[4] == [0.0] * 4 + [1.0] * 4
        assert r[5] == [0, 1, 2, 3, 1, 2, 3, 4]
        assert r[6] == 2.0
