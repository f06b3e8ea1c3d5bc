"""Model-YAML parser and layer interpreter: the counterpart of the reference's nn/tasks.py for this path
(``yaml_model_load`` :1069, ``guess_model_scale`` :1090, ``parse_model`` :826, ``BaseModel._predict_once`` :83,
``BaseModel.fuse`` :146, ``DetectionModel`` :260, ``YOLOv10DetectionModel`` :639).

Differences that are deliberate:
  * strides come from the layer table (each module knows its down-sampling), not from a train-mode probe
    forward on the CPU (tasks.py:278-304) - there is no CPU execution path here;
  * ``_predict_once`` plans Concat buffers ahead: a layer whose output feeds a Concat writes straight into its
    channel slice of the concat buffer, so Concat and Upsample+Concat move no extra bytes;
  * layers nobody consumes (layer 27 of the LPC YAML, SURVEY.md finding 3) are skipped;
  * BN folding is total and happens at pack time; ``fuse()`` is idempotent.
"""
import ast
import contextlib
import math
import os
import re
from copy import deepcopy
from pathlib import Path

import torch
import torch.nn as nn
import yaml

from .. import functional as F
from .modules import (C2f, C2fCIB, CBAM, Concat, Conv, LPC, PSA, SCDown, SPPF, Upsample, space_to_depth, v10Detect)
from .modules.base import LpcModule

CFG_DIR = Path(__file__).resolve().parent.parent / "cfg" / "models"

_MODULES = {"Conv": Conv, "C2f": C2f, "C2fCIB": C2fCIB, "SCDown": SCDown, "SPPF": SPPF, "PSA": PSA, "Concat": Concat,
            "v10Detect": v10Detect, "CBAM": CBAM, "space_to_depth": space_to_depth, "LPC": LPC, "nn.Upsample": Upsample}
_CHANNEL_SCALED = {Conv, C2f, C2fCIB, SCDown, SPPF, PSA, LPC}
_REPEAT_INSIDE = {C2f, C2fCIB}


def make_divisible(x, divisor):
    """utils/ops.py:127-141."""
    if isinstance(divisor, torch.Tensor):
        divisor = int(divisor.max())
    return math.ceil(x / divisor) * divisor


def guess_model_scale(model_path):
    """tasks.py:1090-1106."""
    with contextlib.suppress(AttributeError):
        return re.search(r"yolov\d+([nsblmx])", Path(model_path).stem).group(1)
    return ""


def yaml_model_load(path):
    """tasks.py:1069-1087 for v10 files (never 'unified': the scale letter stays in the file name)."""
    path = Path(path)
    cand = [path, CFG_DIR / "v10" / path.name, CFG_DIR / path.name]
    yaml_file = next((c for c in cand if c.is_file()), None)
    if yaml_file is None:
        raise FileNotFoundError(f"'{path}' does not exist (searched {[str(c) for c in cand]})")
    with open(yaml_file, errors="ignore", encoding="utf-8") as f:
        d = yaml.safe_load(f)
    d["scale"] = guess_model_scale(path)
    d["yaml_file"] = str(path)
    return d


def parse_model(d, ch, verbose=False):
    """tasks.py:826-1066 restricted to the module set of cfg/models/v10: returns (nn.Sequential, save)."""
    max_channels = float("inf")
    nc, scales = d.get("nc"), d.get("scales")
    depth, width = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0)
    if scales:
        scale = d.get("scale")
        if not scale:
            scale = tuple(scales.keys())[0]   # tasks.py:834-839 "no model scale passed"
        depth, width, max_channels = scales[scale]
    ch = [ch]
    layers, save, c2 = [], [], ch[-1]
    for i, (f, n, mname, args) in enumerate(d["backbone"] + d["head"]):
        if mname not in _MODULES:
            raise NotImplementedError(f"module '{mname}' is outside the YOLOv10 / LPC inference path")
        m = _MODULES[mname]
        args = [nc if a == "nc" else a for a in args]
        for j, a in enumerate(args):      # tasks.py:852-855: strings such as 'None' / 'nearest'
            if isinstance(a, str):
                with contextlib.suppress(ValueError, SyntaxError):
                    args[j] = ast.literal_eval(a)
        n = n_ = max(round(n * depth), 1) if n > 1 else n
        if m in _CHANNEL_SCALED:
            c1, c2 = ch[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in _REPEAT_INSIDE:
                args.insert(2, n)
                n = 1
        elif m is Concat:
            c2 = sum(ch[x] for x in f)
        elif m is v10Detect:
            args.append([ch[x] for x in f])
        elif m is CBAM:
            args = [ch[f], *args]
            c2 = ch[f]
        elif m is space_to_depth:
            c2 = 4 * ch[f]
        else:
            c2 = ch[f]
        m_ = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        m_.np = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type = i, f, f"lpc_yolo_b200.nn.modules.{m.__name__}"
        if verbose:
            print(f"{i:>3}{str(f):>20}{n_:>3}{m_.np:10.0f}  {m_.type:<45}{str(args):<30}")
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
    return nn.Sequential(*layers), sorted(save)


class BaseModel(LpcModule):
    """tasks.py:48-258 for inference."""

    compute_dtype = torch.bfloat16
    fold_s2d = True      # space_to_depth + C2f.cv1 -> one 2x2 stride-2 conv (exact same arithmetic)
    fold_upsample = os.environ.get("LPC_FOLD_UPSAMPLE", "1") != "0"      # Upsample + Concat + C2f.cv1 without the upsampled tensor

    def forward(self, x, *args, **kwargs):
        return self.predict(x, *args, **kwargs)

    def predict(self, x, profile=False, visualize=False, augment=False, embed=None):
        if augment or visualize or embed or profile:
            raise NotImplementedError("augment / visualize / embed / profile are outside the inference hot path")
        return self._predict_once(x)

    # ---- concat planning ------------------------------------------------------------------------------
    def _plan(self):
        """For every layer: which layers consume it, and (if one of them is a Concat) where it lands."""
        if getattr(self, "_plan_cache", None) is not None:
            return self._plan_cache
        L = list(self.model)
        consumers = {i: [] for i in range(len(L))}
        for m in L:
            for s in ([m.f] if isinstance(m.f, int) else m.f):
                src = m.i - 1 if s == -1 else s
                if src >= 0:
                    consumers[src].append(m.i)
        dest = {}
        for m in L:
            if isinstance(m, Concat):
                off = 0
                for s in m.f:
                    src = m.i - 1 if s == -1 else s
                    c = self._out_ch[src]
                    if src not in dest and not isinstance(L[src], Concat):
                        dest[src] = (m.i, off)
                    off += c
        live = {i for i, c in consumers.items() if c} | {len(L) - 1}
        # space_to_depth whose only consumer is a C2f: fold it into that block's first 1x1 conv
        fold = {}
        for m in L:
            if isinstance(m, space_to_depth) and self.fold_s2d:
                cons = consumers[m.i]
                if len(cons) == 1 and type(L[cons[0]]) is C2f and L[cons[0]].f == -1 and m.i not in self.save:
                    fold[m.i] = cons[0]
        # ... and a plain stride-1 3x3 Conv whose only consumer is such a folded space_to_depth runs INSIDE that C2f call
        # (conv -> s2d -> cv1 as one kernel where the C library takes the shape, else the same two launches as before)
        prefold = {}
        if self.front_depth <= 0:
            for i, j in fold.items():
                c = L[i - 1] if i > 0 else None
                if (type(c) is Conv and L[i].f == -1 and c.f == -1 and consumers[i - 1] == [i] and (i - 1) not in self.save
                        and (i - 1) not in dest and i - 1 > 0 and j == i + 1
                        and c.conv.kernel_size == (3, 3) and c.conv.stride == (1, 1) and c.conv.groups == 1):
                    prefold[i - 1] = j
        # nn.Upsample -> Concat([-1, j]) -> C2f: the C2f's first 1x1 conv can read the upsampled half of the Concat buffer from the
        # SMALL map (lpc_conv1x1_up2cat_tc), so the Upsample launch and its 4x larger copy disappear.  upfold[i] = index of the C2f.
        upfold = {}
        if self.front_depth <= 0 and self.fold_upsample:
            for m in L:
                i = m.i
                if (isinstance(m, Upsample) and m.f == -1 and i + 2 < len(L) and isinstance(L[i + 1], Concat) and isinstance(L[i + 1].f, list)
                        and len(L[i + 1].f) == 2 and L[i + 1].f[0] == -1 and L[i + 1].f[1] not in (-1, i) and isinstance(L[i + 2], C2f)     # (C2fCIB shares C2f.forward)
                        and L[i + 2].f == -1 and consumers[i] == [i + 1] and consumers[i + 1] == [i + 2] and i not in self.save
                        and (i + 1) not in self.save and (i + 1) not in dest and dest.get(i) == (i + 1, 0) and (i + 1) not in fold
                        and L[i + 1].f[1] in dest and dest[L[i + 1].f[1]][0] == i + 1):
                    upfold[i] = i + 2
        self._plan_cache = (dest, live, fold, prefold, upfold)
        return self._plan_cache

    def stem_u8_supported(self, src):
        """True when ``src`` (uint8 HWC device images [B,H,W,3]) can enter the network without a packed copy: bf16 mode, layer 0
        a plain stride-2 3x3 stem Conv feeding only layer 1, a shape lpc_stem_conv_u8 takes."""
        if self.compute_dtype != torch.bfloat16 or not (torch.is_tensor(src) and src.is_cuda and src.dtype == torch.uint8):
            return False
        dest, live, fold, prefold, upfold = self._plan()
        m = self.model[0]
        if type(m) is not Conv or m.f != -1 or 0 in dest or 0 in fold or 0 in prefold or self.front_depth > 0:
            return False
        return m.u8_supported(src, self.compute_dtype)

    def _predict_once(self, x, tail=None):
        """tasks.py:83-111.  ``tail``: optional callable applied instead of the detect head's forward."""
        if not x.is_cuda:
            raise F.LpcError("lpc-yolo_b200 runs on CUDA tensors only (no CPU fallback)")
        if torch.cuda.current_device() != x.device.index:
            # the C library launches on the CURRENT device's streams: a model living on cuda:1 driven from a process whose
            # current device is cuda:0 would launch there with device-1 pointers
            with torch.cuda.device(x.device):
                return self._predict_once(x, tail)
        dest, live, fold, prefold, upfold = self._plan()
        L = list(self.model)
        y, catbuf = [], {}
        start = 0
        up_small = None
        if x.dtype == torch.uint8:
            # uint8 HWC images [B,H,W,3] (BGR, the array-source contract): layer 0 reads them directly - /255, BGR->RGB and the
            # NHWC padding happen inside the stem kernel (``stem_u8_supported`` says when; callers pack otherwise)
            if not self.stem_u8_supported(x):
                raise F.LpcError("uint8 HWC input is only taken where the fused uint8 stem applies (model.stem_u8_supported)")
            x = L[0].forward_u8(x, swap_rb=True)
            y, start = [x if 0 in self.save else None], 1
        elif x.dtype not in (torch.bfloat16, torch.float32) or not F.is_nhwc_view(x) or x.dtype != self.compute_dtype:
            x = F.pack_input(x.float().contiguous(), self.compute_dtype) if x.shape[1] <= 4 else F.as_act(x, self.compute_dtype)
        front = self._front(x, L, dest, fold) if start == 0 else None
        if front is not None:
            x, start = front
            y = [None] * start
        for m in L[start:]:
            if m.i not in live:
                y.append(None)
                continue
            if m.i in fold or m.i in prefold:      # skipped: the consumer reads our INPUT (s2d fold / conv + s2d fusion)
                y.append(None)
                continue
            if m.i in upfold and (m.i + 1) in catbuf:      # Upsample folded into the C2f two layers on: x stays the SMALL map
                up_small = x
                y.append(None)
                continue
            if m.i - 1 in upfold and up_small is not None:        # its Concat: the buffer the skip source already wrote into
                x = catbuf[m.i]
                y.append(None)
                continue
            extra = {}
            if m.i - 2 in upfold and up_small is not None:        # the C2f: cv1 reads cat[upsample2x(up_small), skip]
                extra, up_small = {"up": up_small}, None
            if m.i - 1 in fold and fold[m.i - 1] == m.i:
                x = m(x, s2d=True, pre=L[m.i - 2] if prefold.get(m.i - 2) == m.i else None)
                y.append(x if m.i in self.save else None)
                continue
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            if m.i in dest and hasattr(m, "out_shape"):
                j, off = dest[m.i]
                B, c, H, W = m.out_shape(x.shape)
                if j not in catbuf:
                    catbuf[j] = F.new_act(B, self._out_ch[j], H, W, x.dtype, x.device)
                x = m(x, out=catbuf[j][:, off:off + c], **extra)
            elif tail is not None and m is L[-1]:
                x = tail(m, x)
            else:
                x = m(x, **extra)
            y.append(x if m.i in self.save else None)
        return x

    # ---- depth-first front ----------------------------------------------------------------------------------
    # The first layers work on the largest maps: at batch 64 every one of their activations (52 ... 420 MB) is far larger
    # than the 126 MB L2, so each layer streams its input from HBM and its output back.  Running the chain of the first
    # ``front_depth + 1`` layers on a few images at a time keeps a chunk's intermediates in L2 between producer and consumer
    # (only the chain's input and its last output cross HBM), at the price of ``front_chunks`` times as many launches.
    # LPC_FRONT="depth,chunks" (e.g. "3,4") switches it on; off by default until measured faster on a shape.
    front_depth, front_chunks = (lambda e: (int(e.split(",")[0]), int(e.split(",")[1])) if e else (0, 1))(os.environ.get("LPC_FRONT", ""))

    def _front(self, x, L, dest, fold):
        depth, chunks = self.front_depth, self.front_chunks
        B = x.shape[0]
        if depth <= 0 or chunks <= 1 or B < 2 * chunks or depth >= len(L) - 1:
            return None
        for m in L[: depth + 1]:
            if m.f != -1 or m.i in self.save or m.i in dest or isinstance(m, Concat):
                return None
        last = L[depth]
        if last.i in fold or not hasattr(last, "out_shape"):
            return None
        # output shape of the chain
        shp = tuple(x.shape)
        for m in L[: depth + 1]:
            if m.i in fold:
                shp = (shp[0], 4 * shp[1], shp[2] // 2, shp[3] // 2)
                continue
            shp = m.out_shape(shp)
        out = F.new_act(B, shp[1], shp[2], shp[3], x.dtype, x.device)
        for c in range(chunks):
            lo, hi = B * c // chunks, B * (c + 1) // chunks
            xc = x[lo:hi]
            for m in L[: depth + 1]:
                if m.i in fold:
                    continue
                kw = {"out": out[lo:hi]} if m is last else {}
                if m.i - 1 in fold and fold[m.i - 1] == m.i:
                    xc = m(xc, s2d=True, **kw)
                else:
                    xc = m(xc, **kw)
        return out, depth + 1

    def fuse(self, verbose=False):
        """tasks.py:146-182.  BN folding (for EVERY conv) happens when weights are packed; nothing to rewrite,
        and unlike the reference a second call is harmless (SURVEY.md finding 2)."""
        return self

    def is_fused(self, thresh=10):
        return True

    def info(self, detailed=False, verbose=True, imgsz=640):
        n_p = sum(p.numel() for p in self.parameters())
        if verbose:
            print(f"{type(self).__name__}: {len(list(self.modules()))} modules, {n_p} parameters")
        return n_p


class DetectionModel(BaseModel):
    """tasks.py:260-312."""

    def __init__(self, cfg="yolov10n.yaml", ch=3, nc=None, verbose=False):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        ch = self.yaml["ch"] = self.yaml.get("ch", ch)
        if nc and nc != self.yaml["nc"]:
            self.yaml["nc"] = nc
        self.model, self.save = parse_model(deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = self.yaml.get("inplace", True)
        # channel count and cumulative stride of every layer, from the table itself
        self._out_ch, down = [], []
        for m in self.model:
            srcs = [m.f] if isinstance(m.f, int) else m.f
            src = [(m.i - 1 if s == -1 else s) for s in srcs]
            cin = [ch if s < 0 else self._out_ch[s] for s in src]
            din = 1.0 if src[0] < 0 else down[src[0]]
            if hasattr(m, "out_shape") and not isinstance(m, Concat):
                o = m.out_shape((1, cin[0], 256, 256))
                self._out_ch.append(o[1])
                down.append(din * 256 / o[2])
            elif isinstance(m, Concat):
                self._out_ch.append(sum(cin))
                down.append(din)
            else:
                self._out_ch.append(cin[0])
                down.append(din)
        det = self.model[-1]
        if isinstance(det, v10Detect):
            det.stride = torch.tensor([down[s] for s in det.f])
            self.stride = det.stride
            det.bias_init()
        else:
            self.stride = torch.Tensor([32])


class YOLOv10DetectionModel(DetectionModel):
    """tasks.py:639-641 (inference side)."""

    batch_streams = int(os.environ.get("LPC_BATCH_STREAMS", "1"))

    def detect(self, x, max_det=300, clip=True, scale_back=None, streams=None):
        """Fused engine path: images -> [B, max_det, 6] detections (xyxy, score, label), never building y.
        ``scale_back``: optional [B,5] (pad_x, pad_y, gain, orig_w, orig_h) - boxes come back in original-image coordinates.
        ``streams`` > 1: the batch is cut into that many contiguous parts that run the whole network as independent chains on
        parallel streams (parallel branches of a captured graph): images are independent, and while one chain sits in the
        fixed latency of a launch boundary (pipeline fill, last-tile epilogue, dependency resolution) the other chain's
        kernel has the SMs."""
        u8 = x.dtype == torch.uint8              # uint8 HWC images [B,H,W,3] where ``stem_u8_supported`` (else callers pack first)
        hw = (tuple(x.shape[1:3]) if u8 else tuple(x.shape[2:])) if clip else None
        n = streams if streams is not None else self.batch_streams
        B = x.shape[0]
        if n <= 1 or B < 2 * n:
            return self._predict_once(x, tail=lambda m, feats: m.detections(feats, max_det, hw, scale_back))
        if not u8 and (x.dtype != self.compute_dtype or not F.is_nhwc_view(x)):
            x = F.pack_input(x.float().contiguous(), self.compute_dtype) if x.shape[1] <= 4 else F.as_act(x, self.compute_dtype)
        out = torch.empty((B, max_det, 6), dtype=torch.float32, device=x.device)
        bounds = [(B * i // n, B * (i + 1) // n) for i in range(n)]

        def part(lo, hi):
            sb = scale_back[lo:hi] if scale_back is not None else None
            self._predict_once(x[lo:hi], tail=lambda m, feats: m.detections(feats, max_det, hw, sb, out=out[lo:hi]))
        F.fork_join([(lambda lo=lo, hi=hi: part(lo, hi)) for lo, hi in bounds], x.device)
        return out
