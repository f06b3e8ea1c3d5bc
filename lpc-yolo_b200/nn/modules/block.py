"""Drop-in counterparts of the reference's nn/modules/block.py classes on the YOLOv10 / LPC path:
``DFL`` :42, ``SPPF`` :156, ``C2f`` :214, ``Bottleneck`` :325, ``RepVGGDW`` :700, ``CIB`` :735, ``C2fCIB`` :758,
``Attention`` :769, ``PSA`` :797, ``SCDown`` :818, ``space_to_depth`` :4063, the module-level shadow ``Conv``
(Mish) :4914, ``SPCA`` :5725, ``LPC`` :5801.

Quirk reproduced on purpose (SURVEY.md finding 1): inside block.py the name ``Conv`` is re-bound to a Mish
variant with no dilation argument, so every class in this file activates with Mish; only YAML-level
``Conv`` layers and the detect head (which import conv.Conv) use SiLU.

chunk / split / cat never move data here: a block allocates its concat buffer once and its convs read and
write channel slices of it (NHWC pitch > C).
"""
import torch
import torch.nn as nn

from ... import functional as F
from ... import pack
from ..._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU
from .base import LpcModule, act_code
from .conv import Conv as _SiLUConv

__all__ = ("DFL", "SPPF", "C2f", "Bottleneck", "RepVGGDW", "CIB", "C2fCIB", "Attention", "PSA", "SCDown",
           "space_to_depth", "SPCA", "LPC", "Conv")


def autopad(k, p=None):
    """block.py:4907-4911 (two-argument shadow)."""
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


class Conv(_SiLUConv):
    """The shadow ``block.Conv`` (block.py:4914-4926): Conv2d + BN + **Mish** (Identity when act=False)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        LpcModule.__init__(self)
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.Mish() if act else nn.Identity()


class DFL(LpcModule):
    """Distribution-focal integral (block.py:42-60).  Holds the arange(16) 1x1 conv for state_dict parity;
    the arithmetic is fused into the tail kernels (lpc_v10_decode*)."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return (x.view(b, 4, self.c1, a).transpose(2, 1).float().softmax(1)
                * torch.arange(self.c1, device=x.device, dtype=torch.float32).view(1, -1, 1, 1)).sum(1)


class Bottleneck(LpcModule):
    """block.py:325-340: two convs, shortcut add fused into the second conv's epilogue."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        x = self._in(x)
        return self.cv2(self.cv1(x), out=out, res=x if self.add else None)


class C2f(LpcModule):
    """block.py:214-237."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))

    def forward(self, x, out=None, s2d=False, pre=None, up=None):
        """``s2d=True``: x is the tensor BEFORE a space_to_depth layer; cv1 then runs as a 2x2 stride-2 conv
        (the s2d + 1x1 fold), so the 4x-channel tensor is never written.  ``pre`` (with s2d): the stride-1 Conv layer in
        front of the space_to_depth, x being ITS input: conv -> s2d -> cv1 then is one kernel where the shape is taken.  ``up``: x is a Concat buffer
        cat[upsample2x(up), skip] whose upsampled part was NOT written (nn.Upsample -> Concat -> C2f of the neck)."""
        x = self._in(x)
        B, _, H, W = x.shape
        if s2d:
            H, W = H // 2, W // 2
        c, n = self.c, len(self.m)
        ybuf = F.new_act(B, (2 + n) * c, H, W, x.dtype, x.device)
        if up is not None:
            # x is the Concat buffer whose first channels WOULD hold upsample2x(up): cv1 reads them from ``up`` itself where the
            # C library takes the shape, else the upsampled tensor is materialised into its slice first
            c0 = up.shape[1]
            pk = self.cv1._packed(x, self.cv1._build)
            done = None
            if isinstance(pk, pack.PackedConv) and F.conv1x1_upcat_supported(up, x[:, c0:], pk, F.view_of(ybuf[:, : 2 * c])[1]):
                done = F.conv1x1_upcat(up, x[:, c0:], pk, out=ybuf[:, : 2 * c])
            if done is None:
                F.upsample2x(up, out=x[:, :c0])
                self.cv1(x, out=ybuf[:, : 2 * c])
        elif s2d:
            self.cv1.forward_s2d(x, out=ybuf[:, : 2 * c], pre=pre)
        else:
            self.cv1(x, out=ybuf[:, : 2 * c])
        for i, m in enumerate(self.m):
            m(ybuf[:, (1 + i) * c:(2 + i) * c], out=ybuf[:, (2 + i) * c:(3 + i) * c])
        return self.cv2(ybuf, out=out)

    forward_split = forward

    def out_shape(self, s):
        return (s[0], self.cv2.conv.out_channels, s[2], s[3])


class RepVGGDW(LpcModule):
    """block.py:700-733: SiLU(dw7x7(x) + dw3x3(x)); the two branches are merged on the host."""

    def __init__(self, ed):
        super().__init__()
        self.conv = Conv(ed, ed, 7, 1, 3, g=ed, act=False)
        self.conv1 = Conv(ed, ed, 3, 1, 1, g=ed, act=False)
        self.dim = ed
        self.act = nn.SiLU()

    def forward(self, x, out=None):
        x = self._in(x)
        pk = self._packed(x, lambda dt, dev: pack.pack_repvggdw(self, dt, dev, ACT_SILU))
        return F.dwconv2d(x, pk, out)

    forward_fuse = forward

    def fuse(self):
        """No-op: the merge happens at pack time (the reference's in-place version breaks on a second call)."""
        return self


class CIB(LpcModule):
    """block.py:735-756."""

    def __init__(self, c1, c2, shortcut=True, e=0.5, lk=False):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = nn.Sequential(
            Conv(c1, c1, 3, g=c1),
            Conv(c1, 2 * c_, 1),
            Conv(2 * c_, 2 * c_, 3, g=2 * c_) if not lk else RepVGGDW(2 * c_),
            Conv(2 * c_, c2, 1),
            Conv(c2, c2, 3, g=c2),
        )
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        x = self._in(x)
        y = x
        for m in list(self.cv1)[:-1]:
            y = m(y)
        return self.cv1[-1](y, out=out, res=x if self.add else None)


class C2fCIB(C2f):
    """block.py:758-766."""

    def __init__(self, c1, c2, n=1, shortcut=False, lk=False, g=1, e=0.5):
        super().__init__(c1, c2, n, shortcut, g, e)
        self.m = nn.ModuleList(CIB(self.c, self.c, shortcut, e=1.0, lk=lk) for _ in range(n))


class Attention(LpcModule):
    """block.py:769-795.  The qkv conv's output channels are re-ordered at pack time from the reference's
    per-head [q|k|v] interleave to [all q | all k | all v], which makes V (and hence ``pe``'s input) one
    contiguous channel slice; lpc_psa_attention then fuses QK^T, softmax and PV."""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def _qkv_perm(self):
        nh, kd, hd = self.num_heads, self.key_dim, self.head_dim
        per = 2 * kd + hd
        q = [h * per + i for h in range(nh) for i in range(kd)]
        k = [h * per + kd + i for h in range(nh) for i in range(kd)]
        v = [h * per + 2 * kd + i for h in range(nh) for i in range(hd)]
        return torch.tensor(q + k + v, dtype=torch.long)

    def forward(self, x, out=None, res=None):
        x = self._in(x)
        nh, kd, hd = self.num_heads, self.key_dim, self.head_dim
        pk = self._packed(x, lambda dt, dev: pack.pack_conv_module(self.qkv, dt, dev, ACT_NONE, out_perm=self._qkv_perm()))
        qkv = F.conv2d(x, pk)
        att = F.psa_attention(qkv, nh, kd, hd)
        y = self.pe(qkv[:, 2 * nh * kd:], res=att)          # attention output + positional encoding of V
        return self.proj(y, out=out, res=res)


class PSA(LpcModule):
    """block.py:797-816."""

    def __init__(self, c1, c2, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.attn = Attention(self.c, attn_ratio=0.5, num_heads=self.c // 64)
        self.ffn = nn.Sequential(Conv(self.c, self.c * 2, 1), Conv(self.c * 2, self.c, 1, act=False))

    def forward(self, x, out=None):
        x = self._in(x)
        c = self.c
        ab = self.cv1(x)                       # [a | b]
        b = ab[:, c:]
        b1 = self.attn(b, res=b)               # b + attn(b)
        self.ffn[1](self.ffn[0](b1), out=b, res=b1)   # b <- b1 + ffn(b1), written back next to a
        return self.cv2(ab, out=out)

    def out_shape(self, s):
        return tuple(s)


class SCDown(LpcModule):
    """block.py:818-825: pointwise (Mish) then strided depthwise (no activation)."""

    def __init__(self, c1, c2, k, s):
        super().__init__()
        self.cv1 = Conv(c1, c2, 1, 1)
        self.cv2 = Conv(c2, c2, k=k, s=s, g=c2, act=False)

    def forward(self, x, out=None):
        return self.cv2(self.cv1(self._in(x)), out=out)

    def out_shape(self, s):
        return self.cv2.out_shape(self.cv1.out_shape(s))


class SPPF(LpcModule):
    """block.py:156-176: the three chained 5x5 max-pools are one kernel writing 5/9/13 windows."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        assert k == 5, "the fused pooling kernel implements k=5 (every v10 / LPC YAML)"
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x, out=None):
        x = self._in(x)
        B, _, H, W = x.shape
        c_ = self.cv1.conv.out_channels
        cat = F.new_act(B, 4 * c_, H, W, x.dtype, x.device)
        self.cv1(x, out=cat[:, :c_])
        F.sppf_pool(cat[:, :c_], cat[:, c_:])
        return self.cv2(cat, out=out)

    def out_shape(self, s):
        return (s[0], self.cv2.conv.out_channels, s[2], s[3])


class space_to_depth(LpcModule):
    """block.py:4063-4070."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x, out=None):
        return F.space_to_depth(self._in(x), out)

    def out_shape(self, s):
        return (s[0], 4 * s[1], s[2] // 2, s[3] // 2)


class SPCA(LpcModule):
    """block.py:5725-5749: three dilated depthwise 3x3 -> pointwise 1x1 (bias), gated per channel by
    sigmoid(W2 relu(W1 avgpool(x))), plus x.  Gate and residual ride in the pointwise conv's epilogue."""

    def __init__(self, c):
        super().__init__()
        self.dilated_convs = nn.ModuleList([
            nn.Conv2d(c, c, kernel_size=3, padding=d, dilation=d, groups=c, bias=False) for d in [1, 2, 3]
        ])
        self.pointwise = nn.Conv2d(3 * c, c, kernel_size=1)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.attention = nn.Sequential(
            nn.Conv2d(c, c // 4, 1, bias=False), nn.ReLU(inplace=True), nn.Conv2d(c // 4, c, 1, bias=False), nn.Sigmoid()
        )

    def _build(self, dtype, device):
        c = self.pointwise.out_channels
        dws = [pack.pack_plain_conv(m, dtype, device) for m in self.dilated_convs]
        pw = pack.pack_plain_conv(self.pointwise, dtype, device)
        w1 = self.attention[0].weight.detach().float().view(c // 4, c).to(device).contiguous()
        w2 = self.attention[2].weight.detach().float().view(c, c // 4).to(device).contiguous()
        return dws, pw, w1, w2

    def forward(self, x, out=None):
        x = self._in(x)
        dws, pw, w1, w2 = self._packed(x, self._build)
        B, c, H, W = x.shape
        feats = F.new_act(B, 3 * c, H, W, x.dtype, x.device)
        # the three dilated depthwise convs and the pooled channel gate only read x: four independent small chains
        got = {}
        jobs = [(lambda j=j, pd=pd: F.dwconv2d(x, pd, out=feats[:, j * c:(j + 1) * c])) for j, pd in enumerate(dws)]
        jobs.append(lambda: got.__setitem__("gate", F.pooled_gate(x, w1, None, ACT_RELU, w2, None, ACT_SIGMOID)))
        F.fork_join(jobs, x.device)
        return F.conv2d(feats, pw, out=out, res=x, chan_scale=got["gate"])


class LPC(LpcModule):
    """Light Perception Convolution (block.py:5801-5825)."""

    def __init__(self, c1, c2, k=1, s=1, g=1, act=True):
        super().__init__()
        c_ = c2 // 2
        self.cv1 = Conv(c1, c_, k, s, None, g, act)
        self.cv2 = Conv(c_, c_, 5, 1, None, c_, act)
        self.spca = SPCA(c_ * 2)

    def forward(self, x, out=None):
        x = self._in(x)
        c_ = self.cv1.conv.out_channels
        B, _, Ho, Wo = self.cv1.out_shape(x.shape)
        x2 = F.new_act(B, 2 * c_, Ho, Wo, x.dtype, x.device)
        self.cv1(x, out=x2[:, :c_])
        self.cv2(x2[:, :c_], out=x2[:, c_:])
        return F.channel_deinterleave(self.spca(x2), out)

    def out_shape(self, s):
        o = self.cv1.out_shape(s)
        return (o[0], 2 * o[1], o[2], o[3])
