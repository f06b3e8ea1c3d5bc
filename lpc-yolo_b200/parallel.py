"""Batch sharding across GPUs: images are independent, so rank r takes a contiguous
slice of the batch, runs the whole path locally, and the only exchange is one gather of the [B_r,K,6] detections.
The reference has no multi-GPU inference (utils/torch_utils.py:138-150 maps any device list to cuda:0)."""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous, balanced slices: the first (batch % world) ranks get one extra image."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(dets, batch=None, group=None):
    """all-gather per-rank detections [b_r,K,6] into [B,K,6] in rank (= image) order.  Equal shards use one
    all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed."""
    if not (dist.is_available() and dist.is_initialized()):
        return dets
    world = dist.get_world_size(group)
    if world == 1:
        return dets
    K, C = dets.shape[1], dets.shape[2]
    if batch is None or batch % world == 0:
        out = torch.empty((world * dets.shape[0], K, C), dtype=dets.dtype, device=dets.device)
        dist.all_gather_into_tensor(out, dets.contiguous(), group=group)
        return out
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx, K, C), dtype=dets.dtype, device=dets.device)
    pad[: dets.shape[0]] = dets
    out = torch.empty((world * mx, K, C), dtype=dets.dtype, device=dets.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


def bind_host_to_gpu(index):
    """Pin the calling thread (and, through the first-touch policy, the pinned host buffers it allocates afterwards) to the
    CPU cores next to GPU ``index`` (NVML's ideal CPU affinity for the device).  With one process per GPU on a two-socket
    host this keeps every rank's host-to-device staging traffic on its own socket: all eight ranks reading one node's DRAM
    is what held the end-to-end metric to 5.5x at 8 GPUs while the device-timed metric scaled 7.9x (SCALE_r01).
    Returns the number of CPUs in the mask, or 0 when NVML / affinity control is unavailable (nothing changed)."""
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(int(index))
        N.nvmlDeviceSetCpuAffinity(h)
        import os
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0
