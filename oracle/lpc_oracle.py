"""CPU oracle for the LPC-YOLO / YOLOv10 inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lpc-yolo_b200/`` may import this file; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

What it is: a from-scratch, *functional* restatement of the reference's algorithm for this path, on
plain CPU tensors (``torch`` is used as the ndarray library: ``conv2d``, ``max_pool2d``, ``softmax``,
``topk`` on CPU - the same ATen entry points the reference itself bottoms out in; the reference has no
kernels of its own).  Every function cites the reference file:line it follows
(paths relative to the reference root ``ultralytics/``).

Pinning: the reference's own tests hold no golden vector for this path, so the
oracle is pinned against outputs of the reference itself, produced in the build container by
``oracle/gen_golden.py`` (which imports the unmodified reference from ``/root/reference``) and
committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

The oracle works on a flat ``state_dict`` whose keys and shapes equal the reference's
(``model.<i>...conv.weight`` etc.), so reference weights can be loaded verbatim.
"""
from __future__ import annotations

import math
import os
import zlib
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))
CFG_DIR = os.path.join(os.path.dirname(_HERE), "lpc-yolo_b200", "cfg", "models", "v10")

MODEL_FILES = {
    "yolov10n": "yolov10n.yaml", "yolov10s": "yolov10s.yaml", "yolov10m": "yolov10m.yaml",
    "yolov10b": "yolov10b.yaml", "yolov10l": "yolov10l.yaml", "yolov10x": "yolov10x.yaml",
    "lpc": "yolov10-SPD-Conv-Tiny-CBAM-LPC.yaml",
}

BN_EPS = 1e-3  # utils/torch_utils.py:342-352 (initialize_weights sets eps=1e-3, momentum=0.03)
REG_MAX = 16   # nn/modules/head.py:35


# ----------------------------------------------------------------------------------------------
# YAML -> resolved layer table          (nn/tasks.py:826-1066 parse_model, :1069-1106 yaml loaders)
# ----------------------------------------------------------------------------------------------
def make_divisible(x: float, divisor: int) -> int:
    """utils/ops.py:127-141."""
    return math.ceil(x / divisor) * divisor


def guess_scale(stem: str) -> str:
    """nn/tasks.py:1090-1106: regex ``yolov\\d+([nsblmx])`` on the file stem, '' when it fails."""
    import re
    m = re.search(r"yolov\d+([nsblmx])", stem)
    return m.group(1) if m else ""


@dataclass
class Layer:
    i: int
    f: object            # int or list[int]
    kind: str            # module name
    args: list           # constructor args after channel resolution
    c2: int              # output channels
    down: float = 1.0    # cumulative stride of the output (input pixels per output pixel)


def load_layers(name: str, nc: Optional[int] = None) -> Tuple[List[Layer], List[int], dict]:
    """Resolve a model YAML into constructor calls exactly as ``parse_model`` does."""
    path = name if os.path.exists(name) else os.path.join(CFG_DIR, MODEL_FILES.get(name, name))
    with open(path) as fh:
        d = yaml.safe_load(fh)
    stem = os.path.splitext(os.path.basename(path))[0]
    scale = guess_scale(stem)
    if nc is not None:
        d["nc"] = nc
    nc = d["nc"]
    scales = d.get("scales")
    depth, width, max_ch = 1.0, 1.0, float("inf")
    if scales:
        if not scale:                      # tasks.py:834-839 "no model scale passed" -> first key
            scale = tuple(scales.keys())[0]
        depth, width, max_ch = scales[scale]
    ch = [3]
    down = [1.0]
    layers: List[Layer] = []
    save: List[int] = []
    channel_scaled = {"Conv", "C2f", "C2fCIB", "SCDown", "SPPF", "PSA", "LPC"}
    repeat_inside = {"C2f", "C2fCIB"}
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        args = [nc if a == "nc" else a for a in args]
        n = max(round(n * depth), 1) if n > 1 else n                     # tasks.py:857
        fin = f if isinstance(f, int) else f[0]
        cin = ch[fin]
        dn = down[fin]
        if m in channel_scaled:
            c1, c2 = cin, args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_ch) * width, 8)          # tasks.py:900
            args = [c1, c2, *args[1:]]
            if m in repeat_inside:
                args.insert(2, n)                                        # tasks.py:912-915
                n = 1
            if m == "Conv":
                dn = dn * (args[3] if len(args) > 3 else 1)
            elif m == "SCDown":
                dn = dn * args[3]
            elif m == "LPC":
                dn = dn * (args[3] if len(args) > 3 else 1)
        elif m == "Concat":
            c2 = sum(ch[x] for x in f)                                   # tasks.py:928-929
        elif m == "v10Detect":
            args = [*args, [ch[x] for x in f]]                           # tasks.py:930-931
            c2 = cin
        elif m == "CBAM":
            args = [cin, *args]                                          # tasks.py:1008-1011
            c2 = cin
        elif m == "space_to_depth":
            c2 = 4 * cin                                                 # tasks.py:1017-1018
            dn = dn * 2
        elif m == "nn.Upsample":
            c2 = cin
            dn = dn / args[1]
        else:
            raise ValueError(f"module {m} is outside the hot path (SURVEY.md section 2)")
        assert n == 1, "no v10/LPC YAML repeats a non-C2f module"
        layers.append(Layer(i, f, m, args, c2, dn))
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)  # tasks.py:1061
        if i == 0:
            ch, down = [], []
        ch.append(c2)
        down.append(dn)
    det = layers[-1]
    strides = [down[x] for x in det.f]
    return layers, sorted(save), {"nc": nc, "scale": scale, "strides": strides, "stem": stem}


# ----------------------------------------------------------------------------------------------
# parameter inventory (constructor restatements; key order = reference state_dict order)
# ----------------------------------------------------------------------------------------------
def _p_conv(P, pre, c1, c2, k=1, g=1):
    """conv.py:41-46 / block.py:4916-4920: Conv2d(bias=False) + BatchNorm2d."""
    P[pre + ".conv.weight"] = (c2, c1 // g, k, k)
    P[pre + ".bn.weight"] = (c2,)
    P[pre + ".bn.bias"] = (c2,)
    P[pre + ".bn.running_mean"] = (c2,)
    P[pre + ".bn.running_var"] = (c2,)
    P[pre + ".bn.num_batches_tracked"] = ()


def _p_c2f(P, pre, c1, c2, n, cib=False, lk=False):
    """block.py:217-225 (C2f), :761-766 (C2fCIB), :738-750 (CIB), :701-706 (RepVGGDW)."""
    c = int(c2 * 0.5)
    _p_conv(P, pre + ".cv1", c1, 2 * c, 1)
    _p_conv(P, pre + ".cv2", (2 + n) * c, c2, 1)
    for j in range(n):
        q = f"{pre}.m.{j}"
        if not cib:
            _p_conv(P, q + ".cv1", c, c, 3)
            _p_conv(P, q + ".cv2", c, c, 3)
        else:
            _p_conv(P, q + ".cv1.0", c, c, 3, g=c)
            _p_conv(P, q + ".cv1.1", c, 2 * c, 1)
            if lk:
                _p_conv(P, q + ".cv1.2.conv", 2 * c, 2 * c, 7, g=2 * c)
                _p_conv(P, q + ".cv1.2.conv1", 2 * c, 2 * c, 3, g=2 * c)
            else:
                _p_conv(P, q + ".cv1.2", 2 * c, 2 * c, 3, g=2 * c)
            _p_conv(P, q + ".cv1.3", 2 * c, c, 1)
            _p_conv(P, q + ".cv1.4", c, c, 3, g=c)


def psa_dims(c1: int) -> Tuple[int, int, int, int]:
    """block.py:770-778, :799-806 -> (c, heads, key_dim, head_dim)."""
    c = int(c1 * 0.5)
    heads = c // 64
    hd = c // heads
    kd = int(hd * 0.5)
    return c, heads, kd, hd


def detect_dims(nc: int, ch: Sequence[int]) -> Tuple[int, int]:
    """head.py:37 (c2 box width), head.py:503 (c3 cls width)."""
    return max(16, ch[0] // 4, REG_MAX * 4), max(ch[0], min(nc, 100))


def _p_detect(P, pre, nc, ch):
    """head.py:30-43 Detect.__init__ then :501-509 v10Detect.__init__ (cv3 replaced, one2one deep copies)."""
    c2, c3 = detect_dims(nc, ch)

    def box(tag):
        for l, x in enumerate(ch):
            _p_conv(P, f"{pre}.{tag}.{l}.0", x, c2, 3)
            _p_conv(P, f"{pre}.{tag}.{l}.1", c2, c2, 3)
            P[f"{pre}.{tag}.{l}.2.weight"] = (4 * REG_MAX, c2, 1, 1)
            P[f"{pre}.{tag}.{l}.2.bias"] = (4 * REG_MAX,)

    def cls(tag):
        for l, x in enumerate(ch):
            _p_conv(P, f"{pre}.{tag}.{l}.0.0", x, x, 3, g=x)
            _p_conv(P, f"{pre}.{tag}.{l}.0.1", x, c3, 1)
            _p_conv(P, f"{pre}.{tag}.{l}.1.0", c3, c3, 3, g=c3)
            _p_conv(P, f"{pre}.{tag}.{l}.1.1", c3, c3, 1)
            P[f"{pre}.{tag}.{l}.2.weight"] = (nc, c3, 1, 1)
            P[f"{pre}.{tag}.{l}.2.bias"] = (nc,)

    box("cv2"); cls("cv3")
    P[f"{pre}.dfl.conv.weight"] = (1, REG_MAX, 1, 1)
    box("one2one_cv2"); cls("one2one_cv3")


def param_shapes(layers: List[Layer]) -> "OrderedDict[str, tuple]":
    """Keys/shapes of the reference ``state_dict`` for this layer table."""
    P: "OrderedDict[str, tuple]" = OrderedDict()
    for L in layers:
        pre = f"model.{L.i}"
        a = L.args
        if L.kind == "Conv":
            _p_conv(P, pre, a[0], a[1], a[2] if len(a) > 2 else 1)
        elif L.kind == "C2f":
            _p_c2f(P, pre, a[0], a[1], a[2])
        elif L.kind == "C2fCIB":
            _p_c2f(P, pre, a[0], a[1], a[2], cib=True, lk=(a[4] if len(a) > 4 else False))
        elif L.kind == "SCDown":
            _p_conv(P, pre + ".cv1", a[0], a[1], 1)
            _p_conv(P, pre + ".cv2", a[1], a[1], a[2], g=a[1])
        elif L.kind == "SPPF":
            _p_conv(P, pre + ".cv1", a[0], a[0] // 2, 1)
            _p_conv(P, pre + ".cv2", a[0] // 2 * 4, a[1], 1)
        elif L.kind == "PSA":
            c, heads, kd, hd = psa_dims(a[0])
            _p_conv(P, pre + ".cv1", a[0], 2 * c, 1)
            _p_conv(P, pre + ".cv2", 2 * c, a[0], 1)
            _p_conv(P, pre + ".attn.qkv", c, c + 2 * kd * heads, 1)
            _p_conv(P, pre + ".attn.proj", c, c, 1)
            _p_conv(P, pre + ".attn.pe", c, c, 3, g=c)
            _p_conv(P, pre + ".ffn.0", c, 2 * c, 1)
            _p_conv(P, pre + ".ffn.1", 2 * c, c, 1)
        elif L.kind == "CBAM":                                   # conv.py:278-320
            c = a[0]
            P[pre + ".channel_attention.fc.weight"] = (c, c, 1, 1)
            P[pre + ".channel_attention.fc.bias"] = (c,)
            P[pre + ".spatial_attention.cv1.weight"] = (1, 2, a[1], a[1])
        elif L.kind == "LPC":                                    # block.py:5801-5809, SPCA :5725-5741
            c_ = a[1] // 2
            _p_conv(P, pre + ".cv1", a[0], c_, a[2])
            _p_conv(P, pre + ".cv2", c_, c_, 5, g=c_)
            c = 2 * c_
            for j in range(3):
                P[f"{pre}.spca.dilated_convs.{j}.weight"] = (c, 1, 3, 3)
            P[pre + ".spca.pointwise.weight"] = (c, 3 * c, 1, 1)
            P[pre + ".spca.pointwise.bias"] = (c,)
            P[pre + ".spca.attention.0.weight"] = (c // 4, c, 1, 1)
            P[pre + ".spca.attention.2.weight"] = (c, c // 4, 1, 1)
        elif L.kind == "v10Detect":
            _p_detect(P, pre, a[0], a[1])
    return P


def synth_state_dict(shapes: "OrderedDict[str, tuple]", seed: int = 0, strides=(8, 16, 32), nc: int = 80):
    """Deterministic synthetic weights keyed by parameter NAME (not by construction order), so the
    reference, the oracle and the CUDA path can all be given bit-identical fp32 weights without sharing
    an RNG stream.  Conv weights ~ U(-a, a) with a = sqrt(3 / fan_in) (unit-gain), BN gamma in
    [0.75, 1.25], beta in [-0.2, 0.2], running stats identity (to be replaced by ``calibrate``), detect
    biases as ``bias_init`` (head.py:88-95, :527-535), DFL weight = arange(16) (block.py:52-54)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shp in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.long)
        elif key.endswith("dfl.conv.weight"):
            t = torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)
        elif ".bn." in key:
            if leaf == "weight":
                t = 0.75 + 0.5 * torch.rand(shp, generator=g)
            elif leaf == "bias":
                t = -0.2 + 0.4 * torch.rand(shp, generator=g)
            elif leaf == "running_mean":
                t = torch.zeros(shp)
            else:
                t = torch.ones(shp)
        elif leaf == "weight":
            fan_in = shp[1] * shp[2] * shp[3]
            a = math.sqrt(3.0 / fan_in)
            t = (torch.rand(shp, generator=g) * 2 - 1) * a
        else:  # conv biases
            parts = key.split(".")
            if "cv2" in parts[2] and parts[-2] == "2":          # box branch: bias 1.0
                t = torch.ones(shp)
            elif "cv3" in parts[2] and parts[-2] == "2":        # cls branch: log(5/nc/(640/s)^2)
                s = strides[int(parts[3])]
                t = torch.full(shp, math.log(5 / nc / (640 / s) ** 2))
                # a small per-class spread so that scores are not constant per level
                t = t + 0.5 * (torch.rand(shp, generator=g) - 0.5)
            else:
                t = 0.1 * (torch.rand(shp, generator=g) - 0.5)
        sd[key] = t
    return sd


# ----------------------------------------------------------------------------------------------
# functional forward
# ----------------------------------------------------------------------------------------------
def mish(x):  # torch.nn.Mish: x * tanh(softplus(x))
    return F.mish(x)


class _Ctx:
    """Forward context: weights + BN mode.  ``calib`` holds running sums when calibrating."""

    def __init__(self, sd, calibrate=False):
        self.sd = sd
        self.calibrate = calibrate
        self.stats: Dict[str, list] = {}

    def bn(self, x, pre):
        """nn.BatchNorm2d forward, eps=1e-3.  Eval: running stats.  Calibrate: batch stats (train
        mode, momentum=None => cumulative average of batch mean / unbiased batch variance)."""
        w, b = self.sd[pre + ".weight"], self.sd[pre + ".bias"]
        if not self.calibrate:
            return F.batch_norm(x, self.sd[pre + ".running_mean"], self.sd[pre + ".running_var"], w, b,
                                False, 0.0, BN_EPS)
        mean = x.mean(dim=(0, 2, 3))
        var_b = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() // x.shape[1]
        self.stats.setdefault(pre, []).append((mean, var_b * n / max(n - 1, 1)))
        return F.batch_norm(x, None, None, w, b, True, 0.0, BN_EPS)

    def conv(self, x, pre, k=1, s=1, g=1, act="silu", p=None, d=1):
        """conv.Conv.forward conv.py:48-50 (act='silu') / block.Conv.forward block.py:4922-4923
        (act='mish'); ``act=None`` is nn.Identity.  autopad conv.py:27-33 / block.py:4907-4911."""
        if p is None:
            p = (d * (k - 1) + 1) // 2
        y = F.conv2d(x, self.sd[pre + ".conv.weight"], None, s, p, d, g)
        y = self.bn(y, pre + ".bn")
        if act == "silu":
            return F.silu(y)
        if act == "mish":
            return mish(y)
        return y


def _bottleneck(cx, x, pre, add):
    """block.py:325-340: two Mish 3x3 convs (+x)."""
    y = cx.conv(cx.conv(x, pre + ".cv1", 3, act="mish"), pre + ".cv2", 3, act="mish")
    return x + y if add else y


def _cib(cx, x, pre, add, lk):
    """block.py:735-756 CIB; RepVGGDW block.py:700-712 (un-fused: SiLU(conv7(x)+conv3(x)))."""
    c = x.shape[1]
    y = cx.conv(x, pre + ".cv1.0", 3, g=c, act="mish")
    y = cx.conv(y, pre + ".cv1.1", 1, act="mish")
    c2 = y.shape[1]
    if lk:
        y = F.silu(cx.conv(y, pre + ".cv1.2.conv", 7, g=c2, act=None, p=3)
                   + cx.conv(y, pre + ".cv1.2.conv1", 3, g=c2, act=None, p=1))
    else:
        y = cx.conv(y, pre + ".cv1.2", 3, g=c2, act="mish")
    y = cx.conv(y, pre + ".cv1.3", 1, act="mish")
    y = cx.conv(y, pre + ".cv1.4", 3, g=y.shape[1], act="mish")
    return x + y if add else y


def _c2f(cx, x, pre, n, shortcut, cib=False, lk=False):
    """block.py:227-231 C2f.forward (chunk(2) -> n bottlenecks on the last chunk -> cat -> cv2)."""
    y = list(cx.conv(x, pre + ".cv1", 1, act="mish").chunk(2, 1))
    for j in range(n):
        q = f"{pre}.m.{j}"
        y.append(_cib(cx, y[-1], q, shortcut, lk) if cib else _bottleneck(cx, y[-1], q, shortcut))
    return cx.conv(torch.cat(y, 1), pre + ".cv2", 1, act="mish")


def _sppf(cx, x, pre):
    """block.py:171-176: three chained MaxPool2d(5,1,2)."""
    x = cx.conv(x, pre + ".cv1", 1, act="mish")
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    y3 = F.max_pool2d(y2, 5, 1, 2)
    return cx.conv(torch.cat((x, y1, y2, y3), 1), pre + ".cv2", 1, act="mish")


def _attention(cx, x, pre, heads, kd, hd):
    """block.py:783-795 Attention.forward."""
    B, C, H, W = x.shape
    N = H * W
    qkv = cx.conv(x, pre + ".qkv", 1, act=None)
    q, k, v = qkv.view(B, heads, 2 * kd + hd, N).split([kd, kd, hd], dim=2)
    attn = (q.transpose(-2, -1) @ k) * (kd ** -0.5)
    attn = attn.softmax(dim=-1)
    y = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + cx.conv(v.reshape(B, C, H, W), pre + ".pe", 3, g=C, act=None)
    return cx.conv(y, pre + ".proj", 1, act=None)


def _psa(cx, x, pre):
    """block.py:812-816 PSA.forward."""
    c, heads, kd, hd = psa_dims(x.shape[1])
    a, b = cx.conv(x, pre + ".cv1", 1, act="mish").split((c, c), dim=1)
    b = b + _attention(cx, b, pre + ".attn", heads, kd, hd)
    b = b + cx.conv(cx.conv(b, pre + ".ffn.0", 1, act="mish"), pre + ".ffn.1", 1, act=None)
    return cx.conv(torch.cat((a, b), 1), pre + ".cv2", 1, act="mish")


def _s2d(x):
    """block.py:4069-4070."""
    return torch.cat([x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]], 1)


def _cbam(cx, x, pre, k):
    """conv.py:286-320: channel attention then spatial attention."""
    sd = cx.sd
    ca = torch.sigmoid(F.conv2d(x.mean((2, 3), keepdim=True), sd[pre + ".channel_attention.fc.weight"],
                                sd[pre + ".channel_attention.fc.bias"]))
    x = x * ca
    m = torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1)
    return x * torch.sigmoid(F.conv2d(m, sd[pre + ".spatial_attention.cv1.weight"], None, 1, k // 2))


def _spca(cx, x, pre):
    """block.py:5743-5749."""
    sd = cx.sd
    c = x.shape[1]
    feats = [F.conv2d(x, sd[f"{pre}.dilated_convs.{j}.weight"], None, 1, d, d, c) for j, d in enumerate((1, 2, 3))]
    spatial = F.conv2d(torch.cat(feats, 1), sd[pre + ".pointwise.weight"], sd[pre + ".pointwise.bias"])
    a = x.mean((2, 3), keepdim=True)
    a = torch.sigmoid(F.conv2d(F.relu(F.conv2d(a, sd[pre + ".attention.0.weight"])), sd[pre + ".attention.2.weight"]))
    return spatial * a + x


def _lpc(cx, x, pre, k, s):
    """block.py:5811-5825: cv1 (k x k stride s, Mish) -> dw5x5 (Mish) -> cat -> SPCA -> channel de-interleave."""
    x1 = cx.conv(x, pre + ".cv1", k, s, act="mish")
    x2 = torch.cat((x1, cx.conv(x1, pre + ".cv2", 5, 1, g=x1.shape[1], act="mish")), 1)
    x2 = _spca(cx, x2, pre + ".spca")
    b, n, h, w = x2.shape
    y = x2.reshape(b * n // 2, 2, h * w).permute(1, 0, 2).reshape(2, -1, n // 2, h, w)
    return torch.cat((y[0], y[1]), 1)


def _head_one2one(cx, feats, pre, nc):
    """head.py:73-77 forward_feat with the one2one branches (head.py:512)."""
    out = []
    for l, x in enumerate(feats):
        p = f"{pre}.one2one_cv2.{l}"
        b = cx.conv(cx.conv(x, p + ".0", 3), p + ".1", 3)
        b = F.conv2d(b, cx.sd[p + ".2.weight"], cx.sd[p + ".2.bias"])
        p = f"{pre}.one2one_cv3.{l}"
        c = cx.conv(cx.conv(x, p + ".0.0", 3, g=x.shape[1]), p + ".0.1", 1)
        c = cx.conv(cx.conv(c, p + ".1.0", 3, g=c.shape[1]), p + ".1.1", 1)
        c = F.conv2d(c, cx.sd[p + ".2.weight"], cx.sd[p + ".2.bias"])
        out.append(torch.cat((b, c), 1))
    return out


def make_anchors(shapes_hw: Sequence[Tuple[int, int]], strides: Sequence[float], dtype=torch.float32):
    """utils/tal.py:294-306 (offset 0.5): anchors [A,2] as (x,y), strides [A,1]; level-major, row-major."""
    pts, st = [], []
    for (h, w), s in zip(shapes_hw, strides):
        sx = torch.arange(w, dtype=dtype) + 0.5
        sy = torch.arange(h, dtype=dtype) + 0.5
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), s, dtype=dtype))
    return torch.cat(pts), torch.cat(st)


def decode(raw: Sequence[torch.Tensor], strides: Sequence[float], nc: int = 80) -> torch.Tensor:
    """Detect.inference head.py:45-71 + DFL block.py:57-60 + dist2bbox(xywh=True) tal.py:309-319.
    raw: per-level [B,64+nc,H,W] -> y [B,4+nc,A] (cx,cy,w,h in pixels, sigmoid scores)."""
    B = raw[0].shape[0]
    no = 4 * REG_MAX + nc
    x_cat = torch.cat([r.reshape(B, no, -1) for r in raw], 2)
    anchors, st = make_anchors([r.shape[2:] for r in raw], strides, raw[0].dtype)
    anchors, st = anchors.transpose(0, 1), st.transpose(0, 1)
    box, cls = x_cat.split((4 * REG_MAX, nc), 1)
    a = box.shape[2]
    prob = box.view(B, 4, REG_MAX, a).transpose(2, 1).softmax(1)
    dist = (prob * torch.arange(REG_MAX, dtype=raw[0].dtype).view(1, REG_MAX, 1, 1)).sum(1)   # 1x1 conv with arange
    lt, rb = dist.split([2, 2], 1)
    x1y1 = anchors.unsqueeze(0) - lt
    x2y2 = anchors.unsqueeze(0) + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


def v10postprocess(preds: torch.Tensor, max_det: int, nc: int = 80):
    """utils/ops.py:851-864, verbatim semantics: preds [B,A,4+nc]."""
    assert 4 + nc == preds.shape[-1]
    boxes, scores = preds.split([4, nc], dim=-1)
    max_scores = scores.amax(dim=-1)
    max_scores, index = torch.topk(max_scores, max_det, dim=-1)
    index = index.unsqueeze(-1)
    boxes = torch.gather(boxes, 1, index.repeat(1, 1, 4))
    scores = torch.gather(scores, 1, index.repeat(1, 1, nc))
    scores, index2 = torch.topk(scores.flatten(1), max_det, dim=-1)
    labels = index2 % nc
    slot = index2 // nc
    boxes = boxes.gather(1, slot.unsqueeze(-1).repeat(1, 1, 4))
    anchor_idx = index.squeeze(-1).gather(1, slot)
    return boxes, scores, labels, anchor_idx


def xywh2xyxy(x):
    """utils/ops.py:402-421."""
    y = torch.empty_like(x)
    dw, dh = x[..., 2] / 2, x[..., 3] / 2
    y[..., 0] = x[..., 0] - dw
    y[..., 1] = x[..., 1] - dh
    y[..., 2] = x[..., 0] + dw
    y[..., 3] = x[..., 1] + dh
    return y


def postprocess(y: torch.Tensor, max_det: int = 300, nc: int = 80, img_hw: Optional[Tuple[int, int]] = None):
    """models/yolov10/predict.py:8-21 up to the ``[B,max_det,6]`` tensor (+ scale_boxes' clamp, ops.py:305-324,
    which is the whole of scale_boxes when source and network shapes agree: gain 1, pad 0).
    Returns (dets [B,K,6] = x1,y1,x2,y2,score,label ; anchor_idx [B,K])."""
    boxes, scores, labels, aidx = v10postprocess(y.transpose(-1, -2), max_det, nc)
    boxes = xywh2xyxy(boxes)
    if img_hw is not None:
        h, w = img_hw
        boxes = torch.stack((boxes[..., 0].clamp(0, w), boxes[..., 1].clamp(0, h),
                             boxes[..., 2].clamp(0, w), boxes[..., 3].clamp(0, h)), -1)
    dets = torch.cat([boxes, scores.unsqueeze(-1), labels.unsqueeze(-1).to(boxes.dtype)], -1)
    return dets, aidx


@dataclass
class OracleModel:
    name: str
    layers: List[Layer]
    save: List[int]
    meta: dict
    sd: "OrderedDict[str, torch.Tensor]" = field(default_factory=OrderedDict)

    @property
    def nc(self):
        return self.meta["nc"]

    @property
    def strides(self):
        return self.meta["strides"]

    # nn/tasks.py:83-111 _predict_once, restricted to what predict() consumes (one2one head only)
    def features(self, x: torch.Tensor, calibrate=False, cx: Optional[_Ctx] = None, run_dead=False):
        cx = cx or _Ctx(self.sd, calibrate)
        y: List[Optional[torch.Tensor]] = []
        det = self.layers[-1]
        needed = set(self.save) | {L.i for L in self.layers}
        if not run_dead:   # layers nobody consumes (LPC layer 27) are skipped; the reference computes and drops them
            used = set()
            for L in self.layers:
                for s in ([L.f] if isinstance(L.f, int) else L.f):
                    used.add(L.i - 1 if s == -1 else s)
            needed = used | {det.i}
        for L in self.layers:
            if L.f != -1:
                x = y[L.f] if isinstance(L.f, int) else [x if j == -1 else y[j] for j in L.f]
            pre = f"model.{L.i}"
            a = L.args
            if L.i not in needed:
                y.append(None)
                continue
            if L.kind == "Conv":
                x = cx.conv(x, pre, a[2] if len(a) > 2 else 1, a[3] if len(a) > 3 else 1)
            elif L.kind == "C2f":
                x = _c2f(cx, x, pre, a[2], a[3] if len(a) > 3 else False)
            elif L.kind == "C2fCIB":
                x = _c2f(cx, x, pre, a[2], a[3] if len(a) > 3 else False, cib=True, lk=(a[4] if len(a) > 4 else False))
            elif L.kind == "SCDown":
                x = cx.conv(cx.conv(x, pre + ".cv1", 1, act="mish"), pre + ".cv2", a[2], a[3], g=a[1], act=None)
            elif L.kind == "SPPF":
                x = _sppf(cx, x, pre)
            elif L.kind == "PSA":
                x = _psa(cx, x, pre)
            elif L.kind == "nn.Upsample":
                x = F.interpolate(x, scale_factor=float(a[1]), mode=a[2])
            elif L.kind == "Concat":
                x = torch.cat(x, 1)
            elif L.kind == "space_to_depth":
                x = _s2d(x)
            elif L.kind == "CBAM":
                x = _cbam(cx, x, pre, a[1])
            elif L.kind == "LPC":
                x = _lpc(cx, x, pre, a[2], a[3])
            elif L.kind == "v10Detect":
                x = _head_one2one(cx, x, pre, a[0])
            y.append(x if L.i in self.save else None)
        return x  # list of 3 raw maps [B,64+nc,H,W]

    def forward(self, x: torch.Tensor):
        """-> (y [B,4+nc,A], raw maps) like ``v10Detect.forward(...)['one2one']`` in eval mode (head.py:516-519)."""
        raw = self.features(x)
        return decode(raw, self.strides, self.nc), raw

    def predict(self, x: torch.Tensor, max_det=300):
        """-> dets [B,max_det,6] (before the conf filter), anchor indices, y, raw."""
        y, raw = self.forward(x)
        dets, aidx = postprocess(y, max_det, self.nc, img_hw=tuple(x.shape[2:]))
        return dets, aidx, y, raw

    def calibrate(self, xs: Sequence[torch.Tensor]):
        """BN calibration: train-mode passes with momentum=None; afterwards
        running_mean/var = average over the passes of batch mean / unbiased batch variance."""
        acc: Dict[str, list] = {}
        for x in xs:
            cx = _Ctx(self.sd, calibrate=True)
            self.features(x, cx=cx)
            for k, v in cx.stats.items():
                acc.setdefault(k, []).extend(v)
        for k, v in acc.items():
            self.sd[k + ".running_mean"] = torch.stack([m for m, _ in v]).mean(0)
            self.sd[k + ".running_var"] = torch.stack([s for _, s in v]).mean(0)
            self.sd[k + ".num_batches_tracked"] = torch.tensor(len(v))
        return self

    def to(self, dtype):
        self.sd = OrderedDict((k, v.to(dtype) if v.is_floating_point() else v) for k, v in self.sd.items())
        return self


def calibration_batches(size: int = 320, n: int = 2, batch: int = 2, seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(batch, 3, size, size, generator=g) for _ in range(n)]


def build(name: str, seed: int = 0, calibrate: bool = True, nc: Optional[int] = None) -> OracleModel:
    layers, save, meta = load_layers(name, nc)
    m = OracleModel(name, layers, save, meta)
    m.sd = synth_state_dict(param_shapes(layers), seed, meta["strides"], meta["nc"])
    if calibrate:
        m.calibrate(calibration_batches())
    return m


def synth_input(batch: int, size: int, seed: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, size, size, generator=g)


def export_output(y: torch.Tensor, max_det: int = 300, nc: int = 80) -> torch.Tensor:
    """v10Detect.forward with export=True (nn/modules/head.py:519-523; engine/exporter.py:227-235 sets the flag): y is
    Detect.inference's [B,4+nc,A]; -> [B,max_det,6] = (cx, cy, w, h, score, label)."""
    boxes, scores, labels, _ = v10postprocess(y.permute(0, 2, 1), max_det, nc)
    return torch.cat([boxes, scores.unsqueeze(-1), labels.unsqueeze(-1).to(boxes.dtype)], dim=-1)


# ---- array-source preprocessing ---------------------------------------------------
def letterbox(img, new_shape=(640, 640), auto=False, stride=32):
    """data/augment.py:684-742 LetterBox.__call__ for one HWC uint8 image (scaleFill=False, scaleup=True, center=True).
    Returns (image, (top, left)).  The resize branch calls cv2.resize exactly as the reference does (:727)."""
    import numpy as np
    shape = img.shape[:2]
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        import cv2
        img = cv2.resize(img, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    img = np.pad(img, ((top, bottom), (left, right), (0, 0)), mode="constant", constant_values=114)   # cv2.copyMakeBorder :729-731
    return img, (top, left)


def preprocess_arrays(ims, imgsz=640, stride=32, pt=True):
    """engine/predictor.py:115-133 + :145-156 for a list of HWC uint8 BGR arrays: LetterBox each (auto = all shapes equal
    and a PyTorch model), stack, BGR->RGB, BHWC->BCHW, float32, /255.  Returns (x[B,3,H,W] float32, (top, left))."""
    import numpy as np
    same = len({x.shape for x in ims}) == 1
    outs = [letterbox(x, imgsz, auto=same and pt, stride=stride) for x in ims]
    im = np.stack([o[0] for o in outs])
    im = np.ascontiguousarray(im[..., ::-1].transpose((0, 3, 1, 2)))
    x = torch.from_numpy(im).float()
    x /= 255
    return x, outs[0][1]


def scale_boxes_unit_gain(boxes, pad_top_left, orig_hw):
    """utils/ops.py:89-124 scale_boxes with gain 1 (+ clip_boxes :305-324): boxes [...,4] xyxy in network coordinates."""
    top, left = pad_top_left
    b = boxes.clone()
    b[..., [0, 2]] -= left
    b[..., [1, 3]] -= top
    b[..., [0, 2]] = b[..., [0, 2]].clamp(0, orig_hw[1])
    b[..., [1, 3]] = b[..., [1, 3]].clamp(0, orig_hw[0])
    return b


def scale_boxes(img1_shape, boxes, img0_shape):
    """utils/ops.py:89-124 scale_boxes (ratio_pad=None, padding=True, xyxy) followed by clip_boxes (:305-324)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    b = boxes.clone()
    b[..., 0] -= pad[0]
    b[..., 1] -= pad[1]
    b[..., 2] -= pad[0]
    b[..., 3] -= pad[1]
    b[..., :4] /= gain
    b[..., [0, 2]] = b[..., [0, 2]].clamp(0, img0_shape[1])
    b[..., [1, 3]] = b[..., [1, 3]].clamp(0, img0_shape[0])
    return b
