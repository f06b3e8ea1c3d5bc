"""Shared machinery of the drop-in modules: lazy per-(dtype, device) weight packing."""
import torch
import torch.nn as nn

from ... import functional as F
from ..._lib import ACT_MISH, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, LpcError

COMPUTE_DTYPES = (torch.bfloat16, torch.float32)

# Bumped whenever any packed-weight cache is dropped (load_state_dict, .to(), invalidate()).  Captured CUDA graphs hold raw
# pointers to packed weights, so the engine keys its graph cache on this counter (engine.py ``_graph_cache``).
_EPOCH = [0]


def weights_epoch():
    return _EPOCH[0]


def _bump(cache):
    if cache:
        _EPOCH[0] += 1
    cache.clear()


def act_code(act):
    """Activation enum from the module's ACTUAL .act object (SURVEY.md finding 1)."""
    if act is None or isinstance(act, nn.Identity):
        return ACT_NONE
    if isinstance(act, nn.SiLU):
        return ACT_SILU
    if isinstance(act, nn.Mish):
        return ACT_MISH
    if isinstance(act, nn.Sigmoid):
        return ACT_SIGMOID
    if isinstance(act, nn.ReLU):
        return ACT_RELU
    raise LpcError(f"activation {type(act).__name__} has no fused epilogue")


class LpcModule(nn.Module):
    """nn.Module whose forward runs liblpcyolo kernels.  Parameters stay fp32 masters with the
    reference's names; packed (BN-folded, re-laid-out) copies are cached per (dtype, device) and dropped
    whenever a state_dict is loaded or the module is moved."""

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_pcache", {})
        self.register_load_state_dict_post_hook(lambda mod, _inc: _bump(mod._pcache))

    def _apply(self, fn, recurse=True):
        _bump(self._pcache)
        return super()._apply(fn, recurse)

    def invalidate(self):
        for m in self.modules():
            if isinstance(m, LpcModule):
                _bump(m._pcache)

    def _packed(self, x, builder):
        key = (x.dtype, x.device)
        pk = self._pcache.get(key)
        if pk is None:
            with torch.no_grad():
                pk = builder(x.dtype, x.device)
            self._pcache[key] = pk
        return pk

    @staticmethod
    def _in(x, like=None):
        """Boundary conversion: accept NCHW-contiguous fp32/bf16 tensors, return an NHWC activation."""
        if not x.is_cuda:
            raise LpcError("lpc-yolo_b200 modules run on CUDA tensors only (no CPU fallback)")
        dtype = x.dtype if x.dtype in COMPUTE_DTYPES else torch.float32
        return F.as_act(x, dtype)
