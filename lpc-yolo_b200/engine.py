"""Predict engine: the counterpart of the reference's ``YOLO(...).predict`` call chain for tensor sources
(engine/model.py:385 ``Model.predict`` -> engine/predictor.py:208 ``stream_inference`` ->
models/yolov10/predict.py:8 ``YOLOv10DetectionPredictor.postprocess`` -> engine/results.py ``Results``).

Kept: argument names and defaults (conf=0.25, max_det=300, classes, half, device, verbose), the three stage
timers reported in ``Results.speed`` (utils/ops.py:18-63 ``Profile``), the predictor callbacks, ``LoadTensor``'s
input validation (data/loaders.py:441-502), ``Results.boxes.data`` = [n,6] (xyxy, conf, cls) on the device.

Changed on purpose: the whole post-processing is one fused kernel pair and ONE host sync per batch (the
reference syncs once per image, predict.py:27); ``Results.orig_img`` is converted to a uint8 HWC array lazily
instead of copying the whole batch to the host on every call (predict.py:29-30).
"""
import time
from collections import defaultdict
from types import SimpleNamespace

import numpy as np
import torch

from . import functional as F
from .nn import checkpoint
from .nn.tasks import YOLOv10DetectionModel

DEFAULTS = dict(conf=0.25, max_det=300, classes=None, half=True, device=None, verbose=False, imgsz=640, batch=1,
                augment=False, visualize=False, embed=None, stream=False, fp32=False)
CALLBACK_EVENTS = ("on_predict_start", "on_predict_batch_start", "on_predict_postprocess_end", "on_predict_batch_end",
                   "on_predict_end")


def _chunk_plan(B):
    """Chunk sizes of a host batch: chunk i+1's H2D copy overlaps chunk i's graph replay.  A small first chunk starts the
    compute early, one large second chunk keeps the kernels efficient (measured on B200, LPC @640 B=64, uint8 source:
    1 chunk 4.7 ms, 2 x 32 4.5 ms, 4 x 16 5.3 ms, 8 x 8 7.6 ms per step; H2D 55 GB/s).  LPC_E2E_CHUNKS=n forces n equal chunks."""
    import os
    plan = os.environ.get("LPC_E2E_PLAN")          # explicit chunk sizes, e.g. "8,24,32" (measurement)
    if plan:
        sizes = [int(v) for v in plan.split(",")]
        if sum(sizes) == B and all(v > 0 for v in sizes):
            return sizes
    env = os.environ.get("LPC_E2E_CHUNKS")
    if env and B % int(env) == 0:
        return [B // int(env)] * int(env)
    if B >= 32 and B % 4 == 0:
        # first chunk ~5/16 of the batch: its replay then lasts as long as the copy of the rest (measured B=64, ms / step with
        # the end-of-round kernels: 12+52 3.85, 16+48 3.69, 20+44 3.55, 24+40 3.64)
        first = max(4, (B * 5 // 16) // 4 * 4)
        return [first, B - first]
    if B >= 8 and B % 2 == 0:
        return [B // 2, B // 2]
    return [B]


def _chunks(B):
    """How many chunks a host batch is cut into so that chunk i+1's H2D copy overlaps chunk i's graph replay.
    LPC_E2E_CHUNKS overrides (measurement)."""
    import os
    env = os.environ.get("LPC_E2E_CHUNKS")
    if env and B % int(env) == 0:
        return int(env)
    return 4 if B % 4 == 0 and B >= 16 else (2 if B % 2 == 0 and B >= 4 else 1)


class Profile:
    """utils/ops.py:18-63: wall-clock stage timer with a device sync on both edges."""

    def __init__(self, device=None):
        self.t, self.dt, self.device = 0.0, 0.0, device

    def __enter__(self):
        self.start = self._time()
        return self

    def __exit__(self, *a):
        self.dt = self._time() - self.start
        self.t += self.dt

    def _time(self):
        if self.device is not None and torch.cuda.is_available():
            torch.cuda.synchronize(self.device)
        return time.time()


class Boxes:
    """engine/results.py Boxes: ``data`` [n,6] = x1,y1,x2,y2,conf,cls."""

    def __init__(self, boxes, orig_shape):
        assert boxes.shape[-1] == 6
        self.data, self.orig_shape = boxes, orig_shape

    xyxy = property(lambda s: s.data[:, :4])
    conf = property(lambda s: s.data[:, 4])
    cls = property(lambda s: s.data[:, 5])

    @property
    def xywh(self):
        b = self.data[:, :4]
        return torch.stack(((b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]), -1)

    def __len__(self):
        return self.data.shape[0]

    def cpu(self):
        return Boxes(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return Boxes(self.data.cpu().numpy(), self.orig_shape)


class Results:
    """engine/results.py Results (detection fields only).  ``boxes`` is built on first access: constructing 2 x B small
    objects per batch was a tenth of a millisecond of host time behind a 2 ms device step."""

    __slots__ = ("_orig", "orig_shape", "_data", "_boxes", "names", "path", "speed")

    def __init__(self, orig_img, path, names, boxes=None, orig_shape=None):
        self._orig = orig_img
        if orig_shape is None:
            orig_shape = tuple(orig_img.shape[-2:]) if torch.is_tensor(orig_img) else orig_img.shape[:2]
        self.orig_shape = orig_shape
        self._data, self._boxes = boxes, None
        self.names, self.path, self.speed = names, path, _NO_SPEED

    @property
    def boxes(self):
        if self._boxes is None and self._data is not None:
            self._boxes = Boxes(self._data, self.orig_shape)
        return self._boxes

    @property
    def orig_img(self):
        """uint8 HWC array as ops.convert_torch2numpy_batch (utils/ops.py:826-836) yields, built on demand."""
        if isinstance(self._orig, tuple):          # (batch tensor, index): sliced only when somebody asks for the image
            self._orig = self._orig[0][self._orig[1]]
        if torch.is_tensor(self._orig):
            self._orig = (self._orig.permute(1, 2, 0).contiguous() * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()
        return self._orig

    def __len__(self):
        return self._data.shape[0] if self._data is not None else 0


_NO_SPEED = {"preprocess": None, "inference": None, "postprocess": None}


def check_tensor_source(im, stride=32):
    """data/loaders.py:463-487 LoadTensor._single_check."""
    s = (f"torch.Tensor inputs should be BCHW i.e. shape(1, 3, 640, 640) divisible by stride {stride}. "
         f"Input shape{tuple(im.shape)} is incompatible.")
    if im.dim() != 4:
        if im.dim() != 3:
            raise ValueError(s)
        im = im.unsqueeze(0)
    if im.shape[2] % stride or im.shape[3] % stride:
        raise ValueError(s)
    return im


class YOLOv10DetectionPredictor:
    """engine/predictor.py BasePredictor + models/yolov10/predict.py for tensor sources."""

    def __init__(self, overrides=None, _callbacks=None):
        a = dict(DEFAULTS)
        a.update(overrides or {})
        self.args = SimpleNamespace(**a)
        self.model = None
        self.device = None
        self.callbacks = _callbacks or defaultdict(list)
        self.results = None
        self.batch = None

    def run_callbacks(self, event):
        for cb in self.callbacks.get(event, []):
            cb(self)

    def add_callback(self, event, func):
        self.callbacks[event].append(func)

    def setup_model(self, model):
        """predictor.py:295-310 / autobackend.py:141-151: move, 'fuse' (pack), pick the compute dtype.
        ``half=True`` selects bf16 (the reference's switch means fp16; it has no bf16 switch)."""
        dev = self.args.device
        if dev is None:            # the process's current CUDA device (one process per GPU under torchrun)
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(dev) if not isinstance(dev, torch.device) else dev
        if self.device.type != "cuda":
            raise F.LpcError("lpc-yolo_b200 predicts on CUDA devices only (no CPU fallback)")
        self.model = model.to(self.device).eval()
        self.model.compute_dtype = torch.float32 if (self.args.fp32 or not self.args.half) else torch.bfloat16
        self.model.fuse()

    def preprocess(self, im):
        """predictor.py:115-133 for tensors: move to the device (no /255 for tensors)."""
        im = check_tensor_source(im)
        if not im.is_cuda:
            im = (im.pin_memory() if not im.is_pinned() else im).to(self.device, non_blocking=True)
        return im.float() if im.dtype != torch.float32 else im

    def inference(self, im):
        """predictor.py:135-142: backbone + neck + one2one head + fused decode/top-k -> [B,K,6]."""
        return self.model.detect(im, self.args.max_det, clip=True)

    # ---- host-source fast path: chunked H2D copies overlapped with CUDA-graph replays ---------------------------
    class _Graphed:
        """Two CUDA graphs of ``model.detect`` on two static input buffers (double buffering)."""

        def __init__(self, model, shape, max_det):
            dev = next(model.parameters()).device
            self.inp = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(2)]
            for t in self.inp:
                t.fill_(0.5)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    model.detect(self.inp[0], max_det, clip=True)
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graphs, self.outs = [], []
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = model.detect(self.inp[i], max_det, clip=True)
                self.graphs.append(g)
                self.outs.append(out)

    def inference_from_host(self, im_host):
        """[B,3,H,W] fp32 host tensor -> [B,K,6] on the device.  The batch is cut into up to four chunks; chunk i+1's
        host-to-device copy (copy stream) overlaps chunk i's graph replay (compute stream)."""
        B = im_host.shape[0]
        n = _chunks(B)
        cb = B // n
        key = (cb, tuple(im_host.shape[1:]), self.args.max_det, self.model.compute_dtype)
        cache = self.__dict__.setdefault("_graphed", {})
        if key not in cache:
            cache[key] = self._Graphed(self.model, (cb, *im_host.shape[1:]), self.args.max_det)
        gd = cache[key]
        if not im_host.is_pinned():
            im_host = im_host.pin_memory()
        cur = torch.cuda.current_stream(self.device)
        cs = self.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=self.device))
        cs.wait_stream(cur)
        preds = torch.empty((B, self.args.max_det, 6), dtype=torch.float32, device=self.device)
        done = []
        for i in range(n):
            b = i & 1
            ev = torch.cuda.Event()
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(done[i - 2])           # the replay that read this buffer has finished
                gd.inp[b].copy_(im_host[i * cb:(i + 1) * cb], non_blocking=True)
                ev.record(cs)
            cur.wait_event(ev)
            gd.graphs[b].replay()
            preds[i * cb:(i + 1) * cb].copy_(gd.outs[b])
            d = torch.cuda.Event()
            d.record(cur)
            done.append(d)
        return preds

    # ---- array sources: uint8 HWC (cv2 / BGR) images, SURVEY.md section 8(f) row 1 ------------------------------
    @staticmethod
    def letterbox_geometry(shape, imgsz, stride=32, auto=True):
        """data/augment.py:700-731 LetterBox for one image shape (h, w) (scaleFill=False, scaleup=True, center=True):
        -> (H, W, top, left, nh, nw): network input size, where the (nh, nw)-resized image lands in it."""
        new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
        new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))         # (w, h)
        dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
        if auto:                                    # minimum rectangle (:716-717)
            dw, dh = dw % stride, dh % stride
        dw, dh = dw / 2, dh / 2                     # center=True (:723-725)
        top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
        left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
        nh, nw = new_unpad[1], new_unpad[0]
        return nh + top + bottom, nw + left + right, top, left, nh, nw

    class _GraphedU8:
        """Two CUDA graphs of pack_u8 + ``model.detect`` on two static uint8 input buffers."""

        def __init__(self, model, cb, hs, ws, geom, max_det):
            dev = next(model.parameters()).device
            H, W, top, left, nh, nw = geom
            self.inp = [torch.full((cb, hs, ws, 3), 114, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.net = [torch.empty((cb, H, W, 4), dtype=model.compute_dtype, device=dev) for _ in range(2)]
            tables = F.resize_tables(hs, ws, nh, nw, dev) if (nh, nw) != (hs, ws) else None

            def run(i):
                if tables is None:                  # LetterBox ratio 1: border + pack only
                    x = F.pack_u8(self.inp[i], model.compute_dtype, top, left, H, W, 114, True, out=self.net[i])
                else:                               # cv2.resize(INTER_LINEAR) fused in front
                    x = F.letterbox_u8(self.inp[i], model.compute_dtype, nh, nw, tables, top, left, H, W, 114, True, out=self.net[i])
                return model.detect(x, max_det, clip=True)

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    run(0)
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graphs, self.outs = [], []
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = run(i)
                self.graphs.append(g)
                self.outs.append(out)

    def inference_from_host_u8(self, im_host):
        """[B,h,w,3] uint8 host tensor (BGR) -> [B,K,6] on the device, in network-input coordinates.  Same chunked
        copy / replay overlap as ``inference_from_host``; the H2D copy moves 3 bytes per pixel instead of 12."""
        B, hs, ws, _ = im_host.shape
        geom = self.letterbox_geometry((hs, ws), self.args.imgsz, int(max(self.model.stride)), auto=True)
        plan = _chunk_plan(B)
        cache = self.__dict__.setdefault("_graphed", {})
        gds = []
        for cb in plan:
            key = ("u8", cb, hs, ws, geom, self.args.max_det, self.model.compute_dtype)
            if key not in cache:
                cache[key] = self._GraphedU8(self.model, cb, hs, ws, geom, self.args.max_det)
            gds.append(cache[key])
        if not im_host.is_pinned():
            im_host = im_host.pin_memory()
        cur = torch.cuda.current_stream(self.device)
        cs = self.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=self.device))
        cs.wait_stream(cur)
        preds = torch.empty((B, self.args.max_det, 6), dtype=torch.float32, device=self.device)
        used = {}                                     # graph object -> buffers already used in this call
        done = {}
        lo = 0
        for cb, gd in zip(plan, gds):
            b = used.get(id(gd), 0) & 1
            used[id(gd)] = used.get(id(gd), 0) + 1
            ev = torch.cuda.Event()
            with torch.cuda.stream(cs):
                if (id(gd), b) in done:
                    cs.wait_event(done[(id(gd), b)])     # the replay that read this buffer has finished
                gd.inp[b].copy_(im_host[lo:lo + cb], non_blocking=True)
                ev.record(cs)
            cur.wait_event(ev)
            gd.graphs[b].replay()
            preds[lo:lo + cb].copy_(gd.outs[b])
            d = torch.cuda.Event()
            d.record(cur)
            done[(id(gd), b)] = d
            lo += cb
        self._pad = (geom[2], geom[3], hs, ws, geom[0], geom[1])
        return preds

    @staticmethod
    def as_u8_batch(source):
        """list of [h,w,3] uint8 arrays (one shape) / one [B,h,w,3] array / uint8 BHWC host tensor -> host tensor."""
        if isinstance(source, (list, tuple)):
            if len({tuple(a.shape) for a in source}) != 1:
                raise NotImplementedError("array sources of different shapes are not batched in this round")
            source = np.stack(source)
        if isinstance(source, np.ndarray):
            if source.ndim == 3:
                source = source[None]
            source = torch.from_numpy(np.ascontiguousarray(source))
        if not (torch.is_tensor(source) and source.dtype == torch.uint8 and source.dim() == 4 and source.shape[3] == 3):
            raise ValueError(f"array sources must be uint8 HWC images, got {getattr(source, 'dtype', type(source))} {tuple(getattr(source, 'shape', ()))}")
        return source.contiguous()

    def scale_back(self, preds):
        """utils/ops.py:89-124 scale_boxes (+ clip_boxes :305-324): network-input coordinates -> original image."""
        top, left, hs, ws, H, W = self._pad
        gain = min(H / hs, W / ws)                                        # :110
        pad_w = round((W - ws * gain) / 2 - 0.1)                          # :111-114
        pad_h = round((H - hs * gain) / 2 - 0.1)
        if gain == 1.0 and pad_w == 0 and pad_h == 0:
            return preds
        preds = preds.clone()
        preds[..., [0, 2]] -= pad_w
        preds[..., [1, 3]] -= pad_h
        preds[..., :4] /= gain
        preds[..., [0, 2]] = preds[..., [0, 2]].clamp_(0, ws)
        preds[..., [1, 3]] = preds[..., [1, 3]].clamp_(0, hs)
        return preds

    def _host_copy(self, preds):
        """Queue ONE device->host copy of the batched detections into a reused pinned buffer (stream order: behind the last
        graph replay).  The copy gives the host path everything it needs - the per-image prefix lengths are counted from
        it instead of by a device reduction plus a second, synchronous read - and it is what ``last_preds_host`` exposes
        (valid until the next call)."""
        buf = self.__dict__.get("_host_preds")
        if buf is None or buf.shape != preds.shape or buf.dtype != preds.dtype:
            buf = torch.empty(preds.shape, dtype=preds.dtype, pin_memory=True)
            self._host_preds = buf
        buf.copy_(preds, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(preds.device))
        return buf, ev

    def postprocess(self, preds, img, orig_imgs, host=None):
        """models/yolov10/predict.py:22-38: confidence / class filter, wrap in Results.  preds are already
        [B,K,6] xyxy (the export-mode contract, head.py:521-523), clipped to the image (scale_boxes is the
        identity + clip when source and network sizes agree)."""
        B = preds.shape[0]
        self.last_preds = preds          # batched [B,K,6] on the device (one D2H gives every detection)
        self.last_preds_host = None
        if host is not None:
            host[1].synchronize()
            self.last_preds_host = host[0]
        K = preds.shape[1]
        if self.args.classes is None:
            # scores are sorted, so a prefix of every image survives: the B prefix lengths come from the host copy of the
            # detections when there is one (else ONE device reduction + one small D2H), ONE split call cuts the flattened
            # [B*K,6] tensor into (kept, dropped) pairs of views
            src = self.last_preds_host if self.last_preds_host is not None else preds
            counts = (src[..., 4] > self.args.conf).sum(1).tolist()
            sizes = [v for n in counts for v in (n, K - n)]
            per_img = preds.reshape(B * K, 6).split(sizes)[0::2]
        else:
            cls = torch.tensor(self.args.classes, device=preds.device, dtype=preds.dtype)
            mask = (preds[..., 4] > self.args.conf) & (preds[..., 5:6] == cls.unsqueeze(0)).any(2)
            per_img = [p[mask[i]] for i, p in enumerate(preds)]
        names = self.model.names
        if torch.is_tensor(orig_imgs):
            shape = tuple(orig_imgs.shape[-2:])
            return [Results((orig_imgs, i), f"image{i}.jpg", names, per_img[i], shape) for i in range(B)]
        return [Results(o, f"image{i}.jpg", names, per_img[i], o.shape[:2]) for i, o in enumerate(orig_imgs)]

    def __call__(self, source):
        if self.model is None:
            raise RuntimeError("setup_model() first")
        self.run_callbacks("on_predict_start")
        profilers = (Profile(self.device), Profile(self.device), Profile(self.device))
        self.batch = source
        self.run_callbacks("on_predict_batch_start")
        with torch.no_grad():
            if not torch.is_tensor(source) or source.dtype == torch.uint8:
                with profilers[0]:
                    im = self.as_u8_batch(source)
                with profilers[1]:
                    preds = self.scale_back(self.inference_from_host_u8(im))
                    host = self._host_copy(preds)
                orig = list(source) if isinstance(source, (list, tuple)) else [a for a in (source if not torch.is_tensor(source) else source.numpy())]
                with profilers[2]:
                    self.results = self.postprocess(preds, im, orig, host=host)
            elif torch.is_tensor(source) and not source.is_cuda and source.dtype == torch.float32 and source.dim() == 4:
                with profilers[0]:
                    im = check_tensor_source(source)
                with profilers[1]:
                    preds = self.inference_from_host(im)
            else:
                with profilers[0]:
                    im = self.preprocess(source)
                with profilers[1]:
                    preds = self.inference(im)
            if torch.is_tensor(source) and source.dtype != torch.uint8:
                with profilers[2]:
                    self.results = self.postprocess(preds, im, im)
        self.run_callbacks("on_predict_postprocess_end")
        n = len(self.results)
        speed = {"preprocess": profilers[0].dt * 1e3 / n, "inference": profilers[1].dt * 1e3 / n,
                 "postprocess": profilers[2].dt * 1e3 / n}         # per-image averages of the batch, as the reference reports
        for r in self.results:
            r.speed = speed
        self.run_callbacks("on_predict_batch_end")
        self.run_callbacks("on_predict_end")
        return self.results


class YOLO:
    """models/yolo/model.py:11 YOLO / models/yolov10/model.py:10 YOLOv10 facade, predict side."""

    def __init__(self, model="yolov10n.yaml", task=None, verbose=False, names=None):
        self.task = task or "detect"
        self.overrides = {}
        self.callbacks = defaultdict(list)
        self.predictor = None
        self.ckpt = self.ckpt_path = None
        if isinstance(model, torch.nn.Module):
            self.model = model
        elif str(model).endswith((".yaml", ".yml")):                       # engine/model.py:195 _new
            self.model = YOLOv10DetectionModel(model, verbose=verbose)
        elif str(model).endswith(".pt"):                                     # engine/model.py:217 _load
            self.model, self.ckpt = checkpoint.attempt_load_one_weight(model)
            self.ckpt_path = self.model.pt_path
        else:
            raise NotImplementedError(f"'{model}': a model YAML, a reference .pt checkpoint or (from_pretrained) a local "
                                      "Hugging Face folder is expected; exported formats are out of scope")
        if names is not None:                                                # models/yolov10/model.py:15-16
            self.model.names = names
        self.names = self.model.names

    @classmethod
    def from_pretrained(cls, folder, **kwargs):
        """models/yolov10/model.py:10 (PyTorchModelHubMixin.from_pretrained), local folders only - no network."""
        model, names, task = checkpoint.from_pretrained_dir(folder)
        return cls(model, task=task, names=names, **kwargs)

    def load(self, weights):
        """engine/model.py:278-298: transfer matching tensors from a checkpoint / state_dict."""
        checkpoint.load_into(self.model, weights)
        return self

    def add_callback(self, event, func):
        self.callbacks[event].append(func)

    def load_state_dict(self, sd, strict=True):
        return self.model.load_state_dict(sd, strict=strict)

    def fuse(self):
        self.model.fuse()
        return self

    def info(self, detailed=False, verbose=True):
        return self.model.info(detailed, verbose)

    def predict(self, source=None, stream=False, predictor=None, **kwargs):
        """engine/model.py:385-441: defaults conf=0.25, then user kwargs."""
        unknown = set(kwargs) - set(DEFAULTS)
        if unknown:  # cfg/__init__.py:302-325 check_dict_alignment raises SyntaxError on unknown keys
            raise SyntaxError(f"'{sorted(unknown)}' are not valid predict() arguments")
        if not torch.is_tensor(source) and not isinstance(source, (np.ndarray, list, tuple)):
            raise NotImplementedError("sources: torch.Tensor [B,3,H,W] in [0,1] (LoadTensor contract) or uint8 HWC BGR arrays "
                                      "(LoadPilAndNumpy contract); files / streams are out of scope")
        args = {**self.overrides, **kwargs}
        if self.predictor is None or args != getattr(self, "_last_args", None):
            self.predictor = (predictor or YOLOv10DetectionPredictor)(overrides=args, _callbacks=self.callbacks)
            self.predictor.setup_model(self.model)
            self._last_args = dict(args)
        return self.predictor(source)

    __call__ = predict


YOLOv10 = YOLO
