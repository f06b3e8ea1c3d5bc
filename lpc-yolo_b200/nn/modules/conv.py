"""Drop-in counterparts of the reference's nn/modules/conv.py classes that the v10 / LPC YAMLs use:
``autopad`` (conv.py:27-33), ``Conv`` (conv.py:36-54), ``ChannelAttention``/``SpatialAttention``/``CBAM``
(conv.py:278-320), ``Concat`` (conv.py:323-333).  Same constructor and forward signatures, same parameter
names; forward runs the sm_100a kernels (lpc_conv2d_*, lpc_dwconv2d, lpc_cbam_*, lpc_copy_channels)."""
import torch
import torch.nn as nn

from ... import functional as F
from ... import pack
from ..._lib import ACT_SIGMOID
from .base import LpcModule, act_code

__all__ = ("autopad", "Conv", "ChannelAttention", "SpatialAttention", "CBAM", "Concat", "Upsample")


def autopad(k, p=None, d=1):
    """'same' padding (conv.py:27-33)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


class Conv(LpcModule):
    """Conv2d(bias=False) + BatchNorm2d + SiLU, fused into one kernel launch."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def _build(self, dtype, device):
        return pack.pack_conv_module(self, dtype, device, act_code(self.act))

    def forward(self, x, out=None, res=None):
        x = self._in(x)
        pk = self._packed(x, self._build)
        if isinstance(pk, pack.PackedDW):
            return F.dwconv2d(x, pk, out, res)
        return F.conv2d(x, pk, out, res)

    forward_fuse = forward  # BN is always folded; kept for API parity (conv.py:52-54)

    def u8_supported(self, src, dtype, out_ld=None):
        """uint8 HWC images [B,H,W,3] as the input of this (stem) conv: does the fused /255 + conv kernel take them?"""
        if dtype != torch.bfloat16 or self.conv.in_channels != 3 or self.conv.groups != 1:
            return False
        pk = self._packed(torch.empty(0, dtype=dtype, device=src.device), self._build)
        return isinstance(pk, pack.PackedConv) and F.stem_u8_supported(src, pk, out_ld)

    def forward_u8(self, src, swap_rb=True, out=None):
        """predictor.preprocess (/255, BGR->RGB, HWC->CHW; engine/predictor.py:115-133) + this conv in one kernel."""
        pk = self._packed(torch.empty(0, dtype=torch.bfloat16, device=src.device), self._build)
        return F.stem_conv_u8(src, pk, swap_rb, out)

    def _build_s2d(self, dtype, device):
        """This 1x1 conv applied to space_to_depth(x) (block.py:4069-4070) == a 2x2 stride-2 conv on x:
        s2d channel q*C + c holds pixel (2y + (q&1), 2x + (q>>1)), so w2[co, c, ky, kx] = w1[co, (ky + 2*kx)*C + c]."""
        w, b = pack.fold_bn(self.conv.weight, self.conv.bias, getattr(self, "bn", None))
        cout, c4 = w.shape[0], w.shape[1]
        c = c4 // 4
        w2 = w.view(cout, 2, 2, c).permute(0, 3, 2, 1).contiguous()      # [co, kx, ky, c] -> [co, c, ky, kx]
        return pack.PackedConv(w2, b, 2, 2, 0, act_code(self.act), dtype, device)

    def _packed_s2d(self, dtype, device):
        key = ("s2d", dtype, device)
        pk = self._pcache.get(key)
        if pk is None:
            with torch.no_grad():
                pk = self._pcache[key] = self._build_s2d(dtype, device)
        return pk

    def forward_s2d(self, x, out=None, pre=None):
        """conv1x1(space_to_depth(x)) without materialising the space_to_depth tensor.  ``pre``: the Conv module that
        produces x from ITS input (then x is that input): 3x3 conv -> s2d -> this 1x1 as one kernel when the shape is taken
        (lpc_conv3x3_s2d_tc), else the two launches."""
        x = self._in(x)
        if pre is not None:
            pk1 = pre._packed(x, pre._build)
            assert self.conv.kernel_size == (1, 1) and self.conv.groups == 1 and self.conv.in_channels == 4 * pre.conv.out_channels
            pk2 = self._packed_s2d(x.dtype, x.device)
            if isinstance(pk1, pack.PackedConv) and F.conv3x3_s2d_supported(x, pk1, pk2, F.view_of(out)[1] if out is not None else None):
                return F.conv3x3_s2d(x, pk1, pk2, out)
            x = pre(x)
        assert self.conv.kernel_size == (1, 1) and self.conv.groups == 1 and self.conv.in_channels == 4 * x.shape[1]
        return F.conv2d(x, self._packed_s2d(x.dtype, x.device), out)

    def out_shape(self, s):
        B, _, H, W = s
        k, st, p, d = self.conv.kernel_size[0], self.conv.stride[0], self.conv.padding[0], self.conv.dilation[0]
        ke = d * (k - 1) + 1
        return (B, self.conv.out_channels, (H + 2 * p - ke) // st + 1, (W + 2 * p - ke) // st + 1)


class ChannelAttention(LpcModule):
    """x * sigmoid(fc(avgpool(x))) (conv.py:278-291): global_avgpool + channel_mlp produce the gate."""

    def __init__(self, channels):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Conv2d(channels, channels, 1, 1, 0, bias=True)
        self.act = nn.Sigmoid()

    def _build(self, dtype, device):
        c = self.fc.out_channels
        return (self.fc.weight.detach().float().view(c, c).to(device).contiguous(), self.fc.bias.detach().float().to(device).contiguous())

    def gate(self, x):
        w, b = self._packed(x, self._build)
        return F.pooled_gate(x, w, b, ACT_SIGMOID)

    def forward(self, x):
        x = self._in(x)
        ca = self.gate(x)
        return (x.float() * ca.view(*ca.shape, 1, 1)).to(x.dtype)  # standalone use only; CBAM fuses this


class SpatialAttention(LpcModule):
    """x * sigmoid(conv_k([mean_c, max_c])) (conv.py:294-307)."""

    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in (3, 7), "kernel size must be 3 or 7"
        self.cv1 = nn.Conv2d(2, 1, kernel_size, padding=3 if kernel_size == 7 else 1, bias=False)
        self.act = nn.Sigmoid()

    def _build(self, dtype, device):
        return self.cv1.weight.detach().float().reshape(-1).to(device).contiguous()

    def forward(self, x, ca=None, out=None):
        x = self._in(x)
        if ca is None:
            ca = torch.ones((x.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
        return F.cbam_spatial(x, ca, self._packed(x, self._build), self.cv1.kernel_size[0], out)


class CBAM(LpcModule):
    """Channel then spatial attention (conv.py:310-320) in three launches over x (pool, stats, apply)."""

    def __init__(self, c1, kernel_size=7):
        super().__init__()
        self.channel_attention = ChannelAttention(c1)
        self.spatial_attention = SpatialAttention(kernel_size)

    def forward(self, x, out=None):
        x = self._in(x)
        return self.spatial_attention(x, ca=self.channel_attention.gate(x), out=out)

    def out_shape(self, s):
        return tuple(s)


class Concat(LpcModule):
    """torch.cat along channels (conv.py:323-333).  When the inputs already are adjacent channel slices of
    one buffer (the model executor arranges that), the buffer itself is returned and nothing is copied."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    @staticmethod
    def _adjacent(xs):
        """The inputs as ONE tensor if they are consecutive channel slices of a single buffer, else None."""
        try:
            p0, ld = F.view_of(xs[0])
            es = xs[0].element_size()
            off = 0
            for t in xs:
                p, l = F.view_of(t)
                if l != ld or t.shape[0] != xs[0].shape[0] or t.shape[2:] != xs[0].shape[2:] or p != p0 + off * es:
                    return None
                if t.untyped_storage().data_ptr() != xs[0].untyped_storage().data_ptr():
                    return None
                off += t.shape[1]
            if off != ld:
                return None
            B, _, H, W = xs[0].shape
            return xs[0].as_strided((B, off, H, W), (H * W * ld, 1, W * ld, ld))
        except (F.LpcError, RuntimeError):
            return None

    def forward(self, x, out=None):
        assert self.d == 1, "only channel concatenation is on the hot path"
        xs = [self._in(t) for t in x]
        if out is None:
            whole = self._adjacent(xs)
            if whole is not None:
                return whole
        B, _, H, W = xs[0].shape
        ctot = sum(t.shape[1] for t in xs)
        if out is None:
            out = F.new_act(B, ctot, H, W, xs[0].dtype, xs[0].device)
        off = 0
        for t in xs:
            F.copy_channels(t, out[:, off:off + t.shape[1]])
            off += t.shape[1]
        return out


class Upsample(LpcModule):
    """nn.Upsample(size=None, scale_factor=2, mode='nearest') as used by every v10 / LPC YAML."""

    def __init__(self, size=None, scale_factor=None, mode="nearest"):
        super().__init__()
        if size is not None or float(scale_factor) != 2.0 or mode != "nearest":
            raise NotImplementedError("only nearest 2x upsampling is on the hot path")
        self.size, self.scale_factor, self.mode = size, scale_factor, mode

    def forward(self, x, out=None):
        return F.upsample2x(self._in(x), out)

    def out_shape(self, s):
        return (s[0], s[1], 2 * s[2], 2 * s[3])
