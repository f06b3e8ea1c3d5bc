"""``make_anchors`` / ``dist2bbox`` with the reference's signatures (utils/tal.py:294-306, :309-319), computed by the
``lpc_make_anchors`` / ``lpc_dist2bbox`` kernels.  The engine's fused tail (``lpc_v10_decode_topk``) derives anchors and
boxes itself and never calls these; they serve code written against the reference's function API."""
import ctypes as C

import torch

from .. import _lib
from ..functional import _fp, _stream


def _need_cuda(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise _lib.LpcError(f"{what} runs on CUDA tensors only (there is no CPU fallback)")


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """-> (anchor_points [A,2] as (x,y), stride_tensor [A,1]) in the dtype of ``feats[0]``; levels in the given order."""
    assert feats is not None
    _need_cuda(feats[0], "make_anchors")
    n = len(strides)
    hw = (C.c_int * (2 * n))(*[int(v) for f in feats[:n] for v in f.shape[2:4]])
    st = (C.c_float * n)(*[float(s) for s in strides])
    A = sum(int(f.shape[2]) * int(f.shape[3]) for f in feats[:n])
    dev = feats[0].device
    pts = torch.empty((A, 2), dtype=torch.float32, device=dev)
    sts = torch.empty((A, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().lpc_make_anchors(n, hw, st, float(grid_cell_offset), _fp(pts), _fp(sts), _stream()), "make_anchors")
    dt = feats[0].dtype
    return (pts, sts) if dt == torch.float32 else (pts.to(dt), sts.to(dt))


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(l,t,r,b) distances along ``dim`` + anchor points (x,y) along ``dim`` (broadcast over leading dims) -> boxes."""
    assert distance.shape[dim] == 4
    _need_cuda(distance, "dist2bbox")
    d = distance.movedim(dim, -1).float().contiguous()
    a = anchor_points.movedim(dim, -1).float()
    while a.dim() > 2 and a.shape[0] == 1:
        a = a[0]
    a = a.expand(*d.shape[:-1], 2).contiguous() if a.dim() > 2 else a.contiguous()
    n, na = d.numel() // 4, a.numel() // 2
    out = torch.empty_like(d)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().lpc_dist2bbox(_fp(d), _fp(a), n, na, int(bool(xywh)), _fp(out), _stream()), "dist2bbox")
    return out.movedim(-1, dim).to(distance.dtype)
