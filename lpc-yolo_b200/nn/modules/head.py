"""Drop-in counterparts of ``Detect`` (reference nn/modules/head.py:21-101) and ``v10Detect`` (:497-535)."""
import copy
import math

import torch
import torch.nn as nn

from ... import functional as F
from ... import pack
from ..._lib import ACT_NONE
from .base import LpcModule
from .block import DFL
from .conv import Conv

__all__ = ("Detect", "v10Detect")


class _Plain1x1(LpcModule):
    """nn.Conv2d(c, n, 1) with bias and no activation, runnable on our kernels while keeping the reference's
    parameter names (``<seq>.2.weight`` / ``.bias``)."""

    def __init__(self, c1, c2):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c2, c1, 1, 1))
        self.bias = nn.Parameter(torch.empty(c2))
        ref = nn.Conv2d(c1, c2, 1)
        with torch.no_grad():
            self.weight.copy_(ref.weight)
            self.bias.copy_(ref.bias)

    def _build(self, dtype, device):
        w, b = pack.fold_bn(self.weight, self.bias, None)
        return pack.PackedConv(w, b, 1, 1, 0, ACT_NONE, dtype, device)

    def forward(self, x, out=None, rowmax=None):
        x = self._in(x)
        return F.conv2d(x, self._packed(x, self._build), out, rowmax=rowmax)


class Detect(LpcModule):
    """YOLOv8 detect head (head.py:21-101): per level a box branch (cv2) and a class branch (cv3)."""

    dynamic = False
    export = False
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), _Plain1x1(c2, 4 * self.reg_max)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), _Plain1x1(c3, self.nc)) for x in ch)
        self.dfl = DFL(self.reg_max)

    def forward_feat(self, x, cv2, cv3):
        """head.py:73-77; the box | class cat is two channel-slice writes into one [B,64+nc,H,W] map."""
        y = []
        for i in range(self.nl):
            xi = self._in(x[i])
            B, _, H, W = xi.shape
            raw = F.new_act(B, self.no, H, W, xi.dtype, xi.device)
            t = xi
            for m in list(cv2[i])[:-1]:
                t = m(t)
            cv2[i][-1](t, out=raw[:, : 4 * self.reg_max])
            t = xi
            for m in list(cv3[i])[:-1]:
                t = m(t)
            cv3[i][-1](t, out=raw[:, 4 * self.reg_max:])
            y.append(raw)
        return y

    def inference(self, x):
        """head.py:45-71 as one kernel: DFL + anchors + dist2bbox(xywh) * stride + sigmoid -> y [B,4+nc,A]."""
        y = F.v10_decode(x, [float(s) for s in self.stride], self.nc)
        return y if self.export else (y, x)

    def forward(self, x):
        y = self.forward_feat(x, self.cv2, self.cv3)
        if self.training:
            return y
        return self.inference(y)

    def bias_init(self):
        """head.py:88-95."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)


class v10Detect(Detect):
    """head.py:497-535.  In eval mode the reference runs both heads and the predictor drops ``one2many``
    (SURVEY.md finding 4); here ``one2many`` is computed only when ``self.compute_one2many`` is set."""

    max_det = 300
    compute_one2many = False

    def __init__(self, nc=80, ch=()):
        super().__init__(nc, ch)
        c3 = max(ch[0], min(self.nc, 100))
        self.cv3 = nn.ModuleList(
            nn.Sequential(nn.Sequential(Conv(x, x, 3, g=x), Conv(x, c3, 1)),
                          nn.Sequential(Conv(c3, c3, 3, g=c3), Conv(c3, c3, 1)),
                          _Plain1x1(c3, self.nc)) for x in ch)
        self.one2one_cv2 = copy.deepcopy(self.cv2)
        self.one2one_cv3 = copy.deepcopy(self.cv3)

    # The six branch chains of the head (3 levels x {box, class}) are independent and, on the 40x40 / 20x20 levels, too
    # small to fill 148 SMs on their own: F.fork_join issues them on side streams (parallel branches of the captured
    # graph).
    def forward_feat(self, x, cv2, cv3, keys=None):
        """keys: optional {"ws": tail workspace, "A": anchors per image}: the last class-branch conv of every level also
        writes the per-anchor max-logit keys there (stage 1 of v10postprocess fused into the conv epilogue)."""
        xs = [self._in(t) for t in x]
        raws, offs, rms = [], [], []
        off = 0
        for xi in xs:
            B, _, H, W = xi.shape
            raws.append(F.new_act(B, self.no, H, W, xi.dtype, xi.device))
            offs.append(off)
            rms.append({"ws": keys["ws"], "A": keys["A"], "off": off} if keys is not None else None)
            off += H * W
        nb = 4 * self.reg_max

        def box(i):
            cv2[i][2](cv2[i][1](cv2[i][0](xs[i])), out=raws[i][:, :nb])

        def cls(i):
            # [dw3x3 + 1x1] -> [dw3x3 + 1x1] -> 1x1 (head.py:504-505).  Fused: two launches (lpc_dwpw_tc: the depthwise output and the
            # first pointwise output never leave the SM) instead of five; shapes the fused kernel declines run link by link.
            x0 = xs[i]
            d0, p0 = cv3[i][0][0]._packed(x0, cv3[i][0][0]._build), cv3[i][0][1]._packed(x0, cv3[i][0][1]._build)
            if F.dwpw_supported(x0, d0, p0):
                t = F.dwpw(x0, d0, p0)
            else:
                t = cv3[i][0][1](cv3[i][0][0](x0))
            d1, p1, p2 = cv3[i][1][0]._packed(t, cv3[i][1][0]._build), cv3[i][1][1]._packed(t, cv3[i][1][1]._build), cv3[i][2]._packed(t, cv3[i][2]._build)
            dst = raws[i][:, nb:]
            if F.dwpw_supported(t, d1, p1, p2, out_ld=F.view_of(dst)[1]):
                F.dwpw(t, d1, p1, p2, out=dst, rowmax=rms[i])
                return
            t = cv3[i][1][1](cv3[i][1][0](t))
            if rms[i] is not None:
                cv3[i][2](t, out=dst, rowmax=rms[i])
            else:
                cv3[i][2](t, out=dst)

        # class branches first: they are the longer chains
        F.fork_join([(lambda fn=fn, i=i: fn(i)) for i in range(self.nl) for fn in (cls, box)], xs[0].device)
        if keys is not None:
            keys["ok"] = all(rm.get("ok", False) for rm in rms)
        return raws

    def forward(self, x):
        one2one = self.forward_feat(x, self.one2one_cv2, self.one2one_cv3)
        one2many = None
        if not self.export and (self.compute_one2many or self.training):
            one2many = Detect.forward(self, x)
        if self.training:
            return {"one2many": one2many, "one2one": one2one}
        if self.export:
            # head.py:519-523: cat(boxes, scores, labels) of ops.v10postprocess on y.  The boxes there are gathered from
            # y as they are, i.e. (cx, cy, w, h) in pixels (decode_bboxes uses dist2bbox(xywh=True), head.py:97-101),
            # unclipped: the fused tail yields xyxy, converted back here.
            d = F.v10_decode_topk(one2one, [float(s) for s in self.stride], self.nc, self.max_det)
            xy1, xy2 = d[..., 0:2], d[..., 2:4]
            return torch.cat(((xy1 + xy2) / 2, xy2 - xy1, d[..., 4:]), -1)
        return {"one2many": one2many, "one2one": self.inference(one2one)}

    def detections(self, x, max_det=None, img_hw=None, scale_back=None, out=None):
        """Engine fast path: raw one2one maps -> fused decode + top-k (+clip, + rescale to the original image when
        ``scale_back`` [B,5] is given) -> [B,K,6]; y is never built."""
        K = max_det or self.max_det
        B = x[0].shape[0]
        A = sum(int(f.shape[2]) * int(f.shape[3]) for f in x)
        keys = {"ws": F.topk_workspace(B, A, K, x[0].device), "A": A}
        one2one = self.forward_feat(x, self.one2one_cv2, self.one2one_cv3, keys=keys)
        return F.v10_decode_topk(one2one, [float(s) for s in self.stride], self.nc, K, img_hw, ws=keys["ws"], keys_ready=keys.get("ok", False),
                                 scale_back=scale_back, out=out)

    def bias_init(self):
        super().bias_init()
        for a, b, s in zip(self.one2one_cv2, self.one2one_cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)
