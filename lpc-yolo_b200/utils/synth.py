"""Deterministic synthetic weights for benchmarks and smoke runs (there are no checkpoints offline).

Name-keyed (crc32 of the parameter name seeds its generator), so any other holder of the same state_dict
keys - e.g. the CPU baseline in bench.py - can be given bit-identical weights by ``state_dict()`` exchange.
Conv weights are unit-gain uniform (a = sqrt(3/fan_in)) times ``gain`` to keep activations O(1) through the
SiLU/Mish stack without a BN calibration pass; BN gamma in [0.75,1.25], beta in [-0.2,0.2], identity stats.
"""
import math
import zlib

import torch


@torch.no_grad()
def init_synthetic(model, seed=0, gain=1.7):
    sd = model.state_dict()
    for key, t in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("num_batches_tracked") or key.endswith("dfl.conv.weight"):
            continue
        if ".bn." in key:
            if leaf == "weight":
                t.copy_(0.75 + 0.5 * torch.rand(t.shape, generator=g))
            elif leaf == "bias":
                t.copy_(-0.2 + 0.4 * torch.rand(t.shape, generator=g))
            elif leaf == "running_mean":
                t.zero_()
            else:
                t.fill_(1.0)
        elif leaf == "weight" and t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            a = math.sqrt(3.0 / fan_in) * gain
            t.copy_((torch.rand(t.shape, generator=g) * 2 - 1) * a)
        elif leaf == "bias" and "cv3" in key:
            t.add_(0.5 * (torch.rand(t.shape, generator=g) - 0.5))      # per-class spread on top of bias_init
    if hasattr(model, "invalidate"):
        model.invalidate()
    return model
