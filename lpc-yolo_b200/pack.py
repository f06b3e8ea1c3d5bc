"""Host-side weight preparation: BatchNorm folding and packing into the kernels' layouts.

Folding follows ``fuse_conv_and_bn`` (reference utils/torch_utils.py:171-198):  W' = diag(g/sqrt(var+eps)) W,
b' = beta - g*mean/sqrt(var+eps) (+ scaled conv bias), eps = 1e-3 (torch_utils.py:342-352) - but it is applied to
EVERY conv+BN pair, including the Mish ``block.Conv`` ones the reference's ``fuse()`` misses (SURVEY.md finding 2),
and RepVGGDW's 3x3 branch is merged into the 7x7 as ``RepVGGDW.fuse`` does (block.py:714-733).  The algebra is done
in fp64 ON THE HOST and rounded once; the device only receives the finished blobs (plain H2D copies: a profiler window
around the first forward sees the network's own kernels, not hundreds of ATen elementwise launches).
"""
import torch

from . import _lib
from ._lib import ACT_NONE


def fold_bn(conv_w, conv_b, bn):
    """-> (W' fp64 [Cout,Cin/g,kh,kw], b' fp64 [Cout]). ``bn`` may be None (plain nn.Conv2d)."""
    w = conv_w.detach().cpu().double()
    cout = w.shape[0]
    b = conv_b.detach().cpu().double() if conv_b is not None else torch.zeros(cout, dtype=torch.float64)
    if bn is None:
        return w, b
    scale = bn.weight.detach().cpu().double() / torch.sqrt(bn.running_var.detach().cpu().double() + bn.eps)
    w = w * scale.view(-1, 1, 1, 1)
    b = b * scale + bn.bias.detach().cpu().double() - bn.running_mean.detach().cpu().double() * scale
    return w, b


class PackedConv:
    """A dense conv ready for lpc_conv2d_direct / lpc_conv2d_tc."""

    def __init__(self, w, b, k, s, p, act, dtype, device, out_perm=None, s2d_fold=False):
        # w: [Cout, Cin, k, k] fp64 folded; optional output-channel permutation (PSA qkv re-ordering)
        w, b = w.cpu(), b.cpu()
        if out_perm is not None:
            w, b = w[out_perm.cpu()], b[out_perm.cpu()]
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.k, self.s, self.p, self.act = k, s, p, act
        w32 = w.float()
        self.bias = b.float().contiguous().to(device)
        self.w_direct = w32.permute(2, 3, 1, 0).reshape(k * k, self.cin, self.cout).contiguous().to(dtype).to(device)
        self.w_tc = None
        self.w_stem = None
        if self.cin == 3 and k == 3 and p == 1 and s in (1, 2) and self.cout % 8 == 0 and self.cout <= 96:
            self.w_stem = w32.permute(2, 3, 1, 0).reshape(27, self.cout).contiguous().to(device)
        if dtype == torch.bfloat16 and self.cin % 16 == 0 and self.cout % 16 == 0 and k in (1, 2, 3):
            kpad = _lib.lib().lpc_conv2d_tc_kpad(self.cin, k)
            if kpad > 0:
                wt = torch.zeros((self.cout, kpad), dtype=torch.float32)
                wt[:, : k * k * self.cin] = w32.permute(0, 2, 3, 1).reshape(self.cout, k * k * self.cin)
                self.w_tc = wt.to(torch.bfloat16).contiguous().to(device)


class PackedDW:
    """A depthwise conv ready for lpc_dwconv2d: weights [k*k][C] fp32."""

    def __init__(self, w, b, k, s, p, d, act, device):
        c = w.shape[0]
        self.c, self.k, self.s, self.p, self.d, self.act = c, k, s, p, d, act
        self.w = w.cpu().float()[:, 0].permute(1, 2, 0).reshape(k * k, c).contiguous().to(device)
        self.bias = b.cpu().float().contiguous().to(device) if b is not None else None


def pack_conv_module(m, dtype, device, act, out_perm=None):
    """m has .conv (nn.Conv2d) and optionally .bn; returns PackedConv or PackedDW by groups."""
    conv = m.conv
    bn = getattr(m, "bn", None)
    w, b = fold_bn(conv.weight, conv.bias, bn)
    k, s, p, d, g = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups
    if g == 1:
        assert d == 1
        return PackedConv(w, b, k, s, p, act, dtype, device, out_perm)
    assert g == conv.in_channels == conv.out_channels, "only depthwise grouping is on the hot path"
    return PackedDW(w, b, k, s, p, d, act, device)


def pack_plain_conv(conv, dtype, device, act=ACT_NONE):
    """Bare nn.Conv2d (detect head's last 1x1, SPCA pointwise / dilated convs)."""
    w, b = fold_bn(conv.weight, conv.bias, None)
    k, s, p, d, g = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0], conv.groups
    if g == 1:
        return PackedConv(w, b, k, s, p, act, dtype, device)
    return PackedDW(w, b if conv.bias is not None else None, k, s, p, d, act, device)


def pack_repvggdw(m, dtype, device, act):
    """RepVGGDW: 7x7 dw + zero-padded 3x3 dw, both BN-folded, summed (block.py:714-733)."""
    w7, b7 = fold_bn(m.conv.conv.weight, None, m.conv.bn)
    w3, b3 = fold_bn(m.conv1.conv.weight, None, m.conv1.bn)
    w = w7 + torch.nn.functional.pad(w3, [2, 2, 2, 2])
    return PackedDW(w, b7 + b3, 7, 1, 3, 1, act, device)
