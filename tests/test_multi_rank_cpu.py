"""world_size-2 gloo test (CPU) of the N>1 host logic: batch sharding + the final gather of detections."""
import importlib
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    par = importlib.import_module("lpc-yolo_b200.parallel")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(batch * 300 * 6, dtype=torch.float32).view(batch, 300, 6)   # stands for the unsharded result
    lo, hi = par.shard_bounds(batch, rank, world)
    got = par.gather_detections(full[lo:hi].clone(), batch)
    q.put((rank, bool(torch.equal(got, full)), (lo, hi)))
    dist.destroy_process_group()


def _run(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    return res


def test_even_and_ragged_shards_gather_in_image_order():
    res = _run(8)
    assert all(ok for _, ok, _ in res) and [b for _, _, b in res] == [(0, 4), (4, 8)]
    res = _run(5)
    assert all(ok for _, ok, _ in res) and [b for _, _, b in res] == [(0, 3), (3, 5)]


def test_shard_bounds_cover_batch(pkg):
    par = importlib.import_module("lpc-yolo_b200.parallel")
    for B in (1, 7, 64, 256):
        for W in (1, 2, 4, 8):
            spans = [par.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
