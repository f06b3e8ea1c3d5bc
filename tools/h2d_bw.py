"""Pinned host -> device copy bandwidth per rank, alone and concurrently (what bounds the end-to-end metric at N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_bw.py
Each phase names the ranks that copy at the same time; the others wait at the barrier.  Output: one JSON line on rank 0."""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
bound = importlib.import_module("lpc-yolo_b200.parallel").bind_host_to_gpu(local) if os.environ.get("H2D_BIND", "1") == "1" else 0
MB = 512
host = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
host.fill_(1)
devbuf = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
phases = [[0]] + [[0, k] for k in (1, 2, 4) if k < world] + ([list(range(0, world, 2))] if world > 2 else []) + [list(range(world))]
out = {}
for ph in phases:
    dist.barrier()
    torch.cuda.synchronize()
    gbs = 0.0
    if rank in ph:
        devbuf.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            devbuf.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = 8 * (MB << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device=dev)
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    out["+".join(map(str, ph))] = [round(v.item(), 1) for v in allv if v.item() > 0]
if rank == 0:
    print(json.dumps({"h2d_gbs_per_rank": out, "cpus_bound_per_rank": bound, "mb_per_copy": MB}))
dist.destroy_process_group()
