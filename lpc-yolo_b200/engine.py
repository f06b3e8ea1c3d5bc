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
_HALF_WARNED = False
_SKIP_H2D = __import__("os").environ.get("LPC_E2E_SKIP_H2D") == "1"      # measurement only: what the copy costs the stream
_SKIP_PREP = __import__("os").environ.get("LPC_E2E_SKIP_PREP") == "1"    # measurement only: ... and the /255 + NHWC pack
CALLBACK_EVENTS = ("on_predict_start", "on_predict_batch_start", "on_predict_postprocess_end", "on_predict_batch_end",
                   "on_predict_end")


def _chunk_plan(B):
    """Chunk sizes of a host batch: chunk i+1's H2D copy overlaps chunk i's graph replay.  A small first chunk starts the
    compute early, one large second chunk keeps the kernels efficient (measured on B200, LPC @640 B=64, uint8 source:
    1 chunk 4.7 ms, 2 x 32 4.5 ms, 4 x 16 5.3 ms, 8 x 8 7.6 ms per step; H2D 55 GB/s).  LPC_E2E_CHUNKS=n forces n equal chunks,
    LPC_E2E_PLAN=a,b,... explicit sizes (measurement)."""
    import os
    if B <= 0:
        return []
    plan = os.environ.get("LPC_E2E_PLAN")
    if plan:
        sizes = [int(v) for v in plan.split(",") if v.strip()]
        if sum(sizes) == B and all(v > 0 for v in sizes):
            return sizes
    env = os.environ.get("LPC_E2E_CHUNKS")
    if env and env.isdigit() and int(env) > 0 and B % int(env) == 0:
        return [B // int(env)] * int(env)
    if B >= 32 and B % 4 == 0:
        # first chunk ~3/8 of the batch: its replay then lasts as long as the copy of the rest (measured B=64, ms / step: round 1
        # kernels 12+52 3.85, 16+48 3.69, 20+44 3.55, 24+40 3.64; end of round 2 - faster kernels, same copy - 16+48 3.68,
        # 20+44 3.34, 24+40 3.30, 28+36 3.39, 32+32 3.49, three chunks 3.48-3.59, one chunk 3.80)
        first = max(4, (B * 3 // 8) // 4 * 4)
        return [first, B - first]
    if B >= 8 and B % 2 == 0:
        return [B // 2, B // 2]
    return [B]


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
    """engine/results.py Results (detection fields only).  Everything per-image is built on first access: ``boxes`` is a view
    of the batch's [B,K,6] detection tensor cut at this image's kept count.  (Slicing 2 x B tensor views and constructing
    B Boxes objects eagerly cost ~0.25 ms of host time per 64-image batch behind a 2 ms device step.)"""

    __slots__ = ("_orig", "orig_shape", "_data", "_boxes", "names", "path", "speed", "_lazy")

    def __init__(self, orig_img, path, names, boxes=None, orig_shape=None, lazy=None):
        self._orig = orig_img
        if orig_shape is None:
            orig_shape = tuple(orig_img.shape[-2:]) if torch.is_tensor(orig_img) else orig_img.shape[:2]
        self.orig_shape = orig_shape
        self._data, self._boxes = boxes, None
        self._lazy = lazy                        # (batch detections [B,K,6], image index, kept count)
        self.names, self.path, self.speed = names, path, _NO_SPEED

    def _rows(self):
        if self._data is None and self._lazy is not None:
            preds, i, n = self._lazy
            self._data = preds[i, :n]
        return self._data

    @property
    def boxes(self):
        if self._boxes is None and self._rows() is not None:
            self._boxes = Boxes(self._data, self.orig_shape)
        return self._boxes

    @property
    def orig_img(self):
        """uint8 HWC array as ops.convert_torch2numpy_batch (utils/ops.py:826-836) yields, built on demand."""
        if isinstance(self._orig, tuple):          # (batch tensor / array, index): sliced only when somebody asks for the image
            self._orig = self._orig[0][self._orig[1]]
        if torch.is_tensor(self._orig):
            if self._orig.dtype == torch.uint8:
                self._orig = self._orig.cpu().numpy()
            else:
                self._orig = (self._orig.permute(1, 2, 0).contiguous() * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()
        return self._orig

    def __len__(self):
        if self._lazy is not None and self._data is None:
            return int(self._lazy[2])
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
        """predictor.py:115-133 for tensors: move to the device (no /255 for float tensors; a uint8 BCHW tensor is the
        0-255 case LoadTensor normalises, data/loaders.py:479-485)."""
        im = check_tensor_source(im)
        if not im.is_cuda:
            im = (im.pin_memory() if not im.is_pinned() else im).to(self.device, non_blocking=True)
        if im.dtype == torch.uint8:
            # tensor / tensor is an IEEE division on the device (tensor / python-scalar multiplies by the rounded reciprocal:
            # 1 ulp off the reference's CPU ``im /= 255`` on some values)
            return im.float() / torch.full((), 255.0, device=im.device)
        return im.float() if im.dtype != torch.float32 else im

    def inference(self, im):
        """predictor.py:135-142: backbone + neck + one2one head + fused decode/top-k -> [B,K,6]."""
        return self.model.detect(im, self.args.max_det, clip=True)

    # ---- host-source fast path: chunked H2D copies overlapped with CUDA-graph replays ---------------------------
    class _Graphed:
        """Two CUDA graphs (double buffering) of ``run(static_input)`` on two static input buffers.  ``stage`` / ``prep``: the
        host data lands in two staging buffers (``stage()``) and ``prep(staging, static_input)`` - launched on the COPY stream,
        outside the graphs - turns it into the graph input (pipelined uint8 streams: the /255 + NHWC pack of batch k+1 then
        runs under batch k's kernels instead of at the head of its own graph)."""

        def __init__(self, dev, make_input, run, stage=None, prep=None):
            self.inp = [make_input() for _ in range(2)]
            self.stage = [stage() for _ in range(2)] if stage is not None else None
            self.prep = prep
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                if prep is not None:
                    for i in range(2):
                        prep(self.stage[i], self.inp[i])
                for _ in range(2):
                    run(self.inp[0])
            torch.cuda.current_stream(dev).wait_stream(side)
            self.graphs, self.outs = [], []
            self.turn, self.done = 0, [None, None]      # next input buffer; "the replay that read buffer b has finished" events
            self.out_done = [None, None]                # "graph b's static output has been copied out" (pipelined streams)
            for i in range(2):
                g = torch.cuda.CUDAGraph()
                # torch.cuda.graph's default capture stream is ONE class-level stream created on whichever device was current
                # first: capturing a cuda:1 model on it switches the current device back to cuda:0 inside the context
                with torch.cuda.graph(g, stream=side):
                    out = run(self.inp[i])
                self.graphs.append(g)
                self.outs.append(out)

    def _graph_cache(self):
        """Captured graphs hold raw pointers to the packed weights: the cache is dropped whenever any packed-weight cache
        was cleared since the capture (load / load_state_dict / .to(): nn/modules/base.py ``weights_epoch``)."""
        from .nn.modules import base
        ep = base.weights_epoch()
        if self.__dict__.get("_graph_epoch") != ep:
            self._graphed = {}
            self._graph_epoch = ep
        return self._graphed

    def _replay_chunks(self, im_host, plan, gds, pipelined=False):
        """Chunk i+1's host-to-device copy (copy stream) overlaps chunk i's graph replay (compute stream).  The two input
        buffers of a graph alternate ACROSS calls and each remembers the event of the last replay that read it, so the copy
        stream never waits for the compute stream as a whole: in a pipelined stream of batches (``stream_inference``) batch
        k+1's copy runs under batch k's replay.  ``pipelined`` also moves the small copies behind a replay (graph output ->
        the batch's detection tensor, later the D2H copy) to a third stream, so that on the compute stream one replay
        follows the other directly; the result then belongs to that stream (``self._out_stream``) until the caller has waited
        for the host copy's event."""
        B = im_host.shape[0]
        if not im_host.is_pinned():
            im_host = im_host.pin_memory()
        cur = torch.cuda.current_stream(self.device)
        cs = self.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=self.device))
        outs = self.__dict__.setdefault("_out_stream", torch.cuda.Stream(device=self.device)) if pipelined else cur
        preds = torch.empty((B, self.args.max_det, 6), dtype=torch.float32, device=self.device)
        if pipelined:
            preds.record_stream(outs)
        lo = 0
        for cb, gd in zip(plan, gds):
            b = gd.turn & 1
            gd.turn += 1
            ev = torch.cuda.Event()
            with torch.cuda.stream(cs):
                if gd.done[b] is not None:
                    cs.wait_event(gd.done[b])            # the replay that read this buffer has finished
                dst = gd.stage[b] if gd.prep is not None else gd.inp[b]
                if not _SKIP_H2D:
                    dst.copy_(im_host[lo:lo + cb], non_blocking=True)
                if gd.prep is not None and not _SKIP_PREP:
                    gd.prep(dst, gd.inp[b])
                ev.record(cs)
            cur.wait_event(ev)
            if gd.out_done[b] is not None:
                cur.wait_event(gd.out_done[b])           # the previous result of this graph has been copied out of its static output
            gd.graphs[b].replay()
            d = torch.cuda.Event()
            d.record(cur)
            gd.done[b] = d
            if pipelined:
                outs.wait_event(d)
                with torch.cuda.stream(outs):
                    preds[lo:lo + cb].copy_(gd.outs[b])
                    od = torch.cuda.Event()
                    od.record(outs)
                gd.out_done[b] = od
            else:
                preds[lo:lo + cb].copy_(gd.outs[b])
            lo += cb
        return preds

    def inference_from_host(self, im_host, pipelined=False):
        """[B,3,H,W] fp32 host tensor -> [B,K,6] on the device, through the same chunk plan as the uint8 path."""
        B = im_host.shape[0]
        plan = [B] if pipelined else _chunk_plan(B)
        cache = self._graph_cache()
        model, K, dev = self.model, self.args.max_det, self.device
        gds = []
        for cb in plan:
            key = ("f32", cb, tuple(im_host.shape[1:]), K, model.compute_dtype)
            if key not in cache:
                shape = (cb, *im_host.shape[1:])
                cache[key] = self._Graphed(dev, lambda shape=shape: torch.full(shape, 0.5, dtype=torch.float32, device=dev),
                                           lambda x: model.detect(x, K, clip=True))
            gds.append(cache[key])
        return self._replay_chunks(im_host, plan, gds, pipelined)

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

    @staticmethod
    def scale_back_row(geom, hs, ws):
        """utils/ops.py:89-124 scale_boxes' constants for an image of shape (hs, ws) letterboxed to ``geom``:
        (pad_x, pad_y, gain, orig_w, orig_h), or None when the rescale is the identity (the tail's clip then is clip_boxes)."""
        H, W = geom[0], geom[1]
        gain = min(H / hs, W / ws)                                        # :110
        pad_w = round((W - ws * gain) / 2 - 0.1)                          # :111-114
        pad_h = round((H - hs * gain) / 2 - 0.1)
        if gain == 1.0 and pad_w == 0 and pad_h == 0:
            return None
        return (float(pad_w), float(pad_h), float(gain), float(ws), float(hs))

    def _prep_u8(self, src, geom, tables, out=None):
        """uint8 HWC device images -> network input (LetterBox resize + border, BGR->RGB, /255, NHWC) in one kernel."""
        H, W, top, left, nh, nw = geom
        dt = self.model.compute_dtype
        if tables is None:                      # LetterBox ratio 1: border + pack only
            return F.pack_u8(src, dt, top, left, H, W, 114, True, out=out)
        return F.letterbox_u8(src, dt, nh, nw, tables, top, left, H, W, 114, True, out=out)

    def _u8_runner(self, cb, hs, ws, geom, split=False):
        """-> run(uint8 [cb,hs,ws,3] device tensor) -> [cb,K,6] in ORIGINAL-image coordinates (rescale fused in the tail).
        ``split``: -> (prep(uint8 images, net buffer [cb,H,W,4]), run(net buffer)) - the pack as a launch of its own."""
        dev, model, K = self.device, self.model, self.args.max_det
        nh, nw = geom[4], geom[5]
        tables = F.resize_tables(hs, ws, nh, nw, dev) if (nh, nw) != (hs, ws) else None
        row = self.scale_back_row(geom, hs, ws)
        scale = torch.tensor([row] * cb, dtype=torch.float32, device=dev) if row is not None else None
        H, W, top, left = geom[:4]
        direct = (tables is None and (top, left, H, W) == (0, 0, hs, ws)
                  and model.stem_u8_supported(torch.empty((1, hs, ws, 3), dtype=torch.uint8, device=dev)))      # (the answer does not depend on the batch)
        if direct:          # no LetterBox border / resize: the stem kernel reads the uint8 image itself (no packed copy of the batch)
            run = lambda inp: model.detect(inp, K, clip=True, scale_back=scale)
            return (None, run) if split else run
        if split:
            return (lambda src, net: self._prep_u8(src, geom, tables, out=net),
                    lambda net: model.detect(net.permute(0, 3, 1, 2)[:, :3], K, clip=True, scale_back=scale))
        return lambda inp: model.detect(self._prep_u8(inp, geom, tables), K, clip=True, scale_back=scale)

    def inference_from_host_u8(self, im_host, pipelined=False):
        """[B,h,w,3] uint8 host tensor (BGR, one shape) -> [B,K,6] on the device in original-image coordinates.  Chunked
        copy / replay overlap; the H2D copy moves 3 bytes per pixel instead of 12.  ``pipelined`` (a stream of batches): one
        whole-batch graph - the overlap then is between consecutive batches, and the second graph's fixed cost goes away."""
        B, hs, ws, _ = im_host.shape
        geom = self.letterbox_geometry((hs, ws), self.args.imgsz, int(max(self.model.stride)), auto=True)
        plan = [B] if pipelined else _chunk_plan(B)
        cache = self._graph_cache()
        dev = self.device
        gds = []
        for cb in plan:
            key = ("u8p" if pipelined else "u8", cb, hs, ws, geom, self.args.max_det, self.model.compute_dtype)
            if key not in cache:
                stage = lambda cb=cb: torch.full((cb, hs, ws, 3), 114, dtype=torch.uint8, device=dev)
                if pipelined:
                    prep, run = self._u8_runner(cb, hs, ws, geom, split=True)
                    net = lambda cb=cb: torch.zeros((cb, geom[0], geom[1], 4), dtype=self.model.compute_dtype, device=dev)
                    cache[key] = self._Graphed(dev, net, run, stage=stage, prep=prep) if prep is not None else self._Graphed(dev, stage, run)
                else:
                    cache[key] = self._Graphed(dev, stage, self._u8_runner(cb, hs, ws, geom))
            gds.append(cache[key])
        return self._replay_chunks(im_host, plan, gds, pipelined)

    def inference_u8_device(self, im):
        """uint8 HWC images already on the device (one shape): no copy, eager launches."""
        B, hs, ws, _ = im.shape
        geom = self.letterbox_geometry((hs, ws), self.args.imgsz, int(max(self.model.stride)), auto=True)
        return self._u8_runner(B, hs, ws, geom)(im.contiguous())

    def inference_mixed_shapes(self, ims):
        """A list of HWC uint8 arrays of DIFFERENT shapes: the reference letterboxes each image to the full imgsz x imgsz
        square (engine/predictor.py:144-156: ``LetterBox(auto=same_shapes and pt)`` - no minimum rectangle when shapes
        differ) and runs them as one batch.  Images are grouped by shape; each group is copied and letterboxed into its
        rows of one network-input buffer, one forward runs over the whole batch, the rescale constants are per image."""
        B = len(ims)
        dev, model, K, imgsz = self.device, self.model, self.args.max_det, self.args.imgsz
        stride = int(max(model.stride))
        groups = {}
        for i, a in enumerate(ims):
            if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3):
                raise ValueError(f"array sources must be uint8 HWC images, got {getattr(a, 'dtype', type(a))} {getattr(a, 'shape', ())}")
            groups.setdefault(a.shape[:2], []).append(i)
        S = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        net = torch.empty((B, S[0], S[1], 4), dtype=model.compute_dtype, device=dev)
        rows = [None] * B
        order = []
        pos = 0
        for (hs, ws), idx in groups.items():
            geom = self.letterbox_geometry((hs, ws), imgsz, stride, auto=False)
            assert (geom[0], geom[1]) == S
            H, W, top, left, nh, nw = geom
            src = torch.from_numpy(np.ascontiguousarray(np.stack([ims[i] for i in idx]))).pin_memory().to(dev, non_blocking=True)
            out = net[pos:pos + len(idx)]
            if (nh, nw) == (hs, ws):
                F.pack_u8(src, model.compute_dtype, top, left, H, W, 114, True, out=out)
            else:
                F.letterbox_u8(src, model.compute_dtype, nh, nw, F.resize_tables(hs, ws, nh, nw, dev), top, left, H, W, 114, True, out=out)
            row = self.scale_back_row(geom, hs, ws) or (0.0, 0.0, 1.0, float(ws), float(hs))
            for j, i in enumerate(idx):
                rows[pos + j] = row
                order.append(i)                   # batch row pos + j holds source image i
            pos += len(idx)
        scale = torch.tensor(rows, dtype=torch.float32, device=dev)
        preds = model.detect(net.permute(0, 3, 1, 2)[:, :3], K, clip=True, scale_back=scale)
        inv = torch.empty(B, dtype=torch.long)
        inv[torch.tensor(order)] = torch.arange(B)
        return preds[inv.to(dev)]

    @staticmethod
    def as_u8_batch(source):
        """list of [h,w,3] uint8 arrays (one shape) / one [B,h,w,3] array / uint8 BHWC tensor -> one tensor."""
        if isinstance(source, (list, tuple)):
            source = np.stack(source)
        if isinstance(source, np.ndarray):
            if source.ndim == 3:
                source = source[None]
            source = torch.from_numpy(np.ascontiguousarray(source))
        if not (torch.is_tensor(source) and source.dtype == torch.uint8 and source.dim() == 4 and source.shape[3] == 3):
            raise ValueError(f"array sources must be uint8 HWC images, got {getattr(source, 'dtype', type(source))} {tuple(getattr(source, 'shape', ()))}")
        return source.contiguous()

    def _host_copy(self, preds, stream=None):
        """Queue ONE device->host copy of the batched detections into a pinned buffer (stream order: behind the last graph
        replay).  The copy gives the host path everything it needs - the per-image prefix lengths are counted from it
        instead of by a device reduction plus a second, synchronous read - and it is what ``last_preds_host`` exposes.  The
        buffers form a ring of three: a copy stays valid while the next two batches are queued (pipelined streams run one
        batch ahead)."""
        ring = self.__dict__.setdefault("_host_ring", [])
        if not ring or ring[0].shape != preds.shape or ring[0].dtype != preds.dtype:
            ring[:] = [torch.empty(preds.shape, dtype=preds.dtype, pin_memory=True) for _ in range(3)]
            self._host_turn = 0
        buf = ring[self._host_turn % 3]
        self._host_turn += 1
        st = stream if stream is not None else torch.cuda.current_stream(preds.device)
        with torch.cuda.stream(st):
            buf.copy_(preds, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        return buf, ev

    def postprocess(self, preds, img, orig_imgs, host=None):
        """models/yolov10/predict.py:22-38: confidence / class filter, wrap in Results.  preds are already
        [B,K,6] xyxy (the export-mode contract, head.py:521-523) in original-image coordinates, clipped (scale_boxes is the
        identity + clip when source and network sizes agree; otherwise it ran inside the tail kernel)."""
        B = preds.shape[0]
        self.last_preds = preds          # batched [B,K,6] on the device (one D2H gives every detection)
        self.last_preds_host = None
        if host is not None:
            host[1].synchronize()
            self.last_preds_host = host[0]
        names = self.model.names
        if torch.is_tensor(orig_imgs) or (isinstance(orig_imgs, np.ndarray) and orig_imgs.ndim == 4):
            hwc = orig_imgs.shape[-1] == 3 and orig_imgs.dtype in (torch.uint8, np.uint8)
            shapes = [tuple(orig_imgs.shape[1:3]) if hwc else tuple(orig_imgs.shape[-2:])] * B
            origs = [(orig_imgs, i) for i in range(B)]
        else:
            shapes = [o.shape[:2] for o in orig_imgs]
            origs = orig_imgs
        if self.args.classes is None:
            # scores are sorted, so a prefix of every image survives: the B prefix lengths come from the host copy of the
            # detections when there is one (else ONE device reduction + one small D2H); each Results cuts its own view lazily
            src = self.last_preds_host if self.last_preds_host is not None else preds
            counts = (src[..., 4] > self.args.conf).sum(1).tolist()
            return [Results(origs[i], f"image{i}.jpg", names, None, shapes[i], lazy=(preds, i, counts[i])) for i in range(B)]
        cls = torch.tensor(self.args.classes, device=preds.device, dtype=preds.dtype)
        mask = (preds[..., 4] > self.args.conf) & (preds[..., 5:6] == cls.unsqueeze(0)).any(2)
        return [Results(origs[i], f"image{i}.jpg", names, p[mask[i]], shapes[i]) for i, p in enumerate(preds)]

    @staticmethod
    def _is_array_source(source):
        """uint8 HWC images (the LoadPilAndNumpy contract): arrays / lists of arrays, or a uint8 tensor whose LAST dim is 3.
        A uint8 tensor shaped [B,3,H,W] is a LoadTensor source (0-255 values, normalised by preprocess)."""
        if isinstance(source, (np.ndarray, list, tuple)):
            return True
        return torch.is_tensor(source) and source.dtype == torch.uint8 and source.dim() in (3, 4) and source.shape[-1] == 3

    def __call__(self, source):
        if self.model is None:
            raise RuntimeError("setup_model() first")
        with torch.cuda.device(self.device):          # the C library launches on the current device's streams
            self.run_callbacks("on_predict_start")
            res = self._finish(self._enqueue(source))
            self.run_callbacks("on_predict_end")
            return res

    @staticmethod
    def source_len(source):
        if isinstance(source, (list, tuple)):
            return len(source)
        if isinstance(source, np.ndarray):
            return source.shape[0] if source.ndim == 4 else 1
        return source.shape[0] if source.dim() == 4 else 1

    def stream_inference(self, source, batch):
        """engine/predictor.py:208-283 ``stream_inference`` for a source of MORE images than one batch: a generator of per-image
        Results, ``batch`` images per step.  The steps are pipelined one batch ahead: batch k+1's host-to-device copy, graph
        replay and result copy are queued before the host waits for batch k's detections, so in steady state the copy engine
        works under the previous batch's kernels and the step time is max(copy, compute) instead of their sum.  The stage
        timers do not synchronise the device here (``Results.speed`` holds host queueing times)."""
        if self.model is None:
            raise RuntimeError("setup_model() first")
        n = self.source_len(source)
        batch = max(1, int(batch))
        if torch.is_tensor(source) and source.dim() == 3 or isinstance(source, np.ndarray) and source.ndim == 3:
            source = source[None]
        import os
        from collections import deque
        depth = max(1, int(os.environ.get("LPC_STREAM_DEPTH", "1")))   # steps queued ahead of the one the host waits for (<= 2:
        depth = min(depth, 2)                                           # three pinned result buffers, two input buffers per graph)
        guard = lambda: torch.cuda.device(self.device)       # entered per call: a generator must not hold the caller's device
        self.run_callbacks("on_predict_start")
        queued = deque()
        for lo in range(0, n, batch):
            with guard():
                queued.append(self._enqueue(source[lo:lo + batch], pipelined=True))
            if len(queued) > depth:
                with guard():
                    res = self._finish(queued.popleft())
                yield from res
        while queued:
            with guard():
                res = self._finish(queued.popleft())
            yield from res
        self.run_callbacks("on_predict_end")

    def _enqueue(self, source, pipelined=False):
        """Queue one batch (copies, launches, the result copy) without waiting for it -> ticket for ``_finish``."""
        tdev = None if pipelined else self.device                  # pipelined: the timers must not drain the device
        profilers = (Profile(tdev), Profile(tdev), Profile(tdev))
        self.batch = source
        self.run_callbacks("on_predict_batch_start")
        host = None
        with torch.no_grad():
            if self._is_array_source(source):
                mixed = isinstance(source, (list, tuple)) and len({tuple(np.shape(a)) for a in source}) > 1
                if isinstance(source, (list, tuple)) and len(source) == 0:
                    raise ValueError("empty source")
                with profilers[0]:
                    im = None if mixed else self.as_u8_batch(source)
                with profilers[1]:
                    if mixed:
                        preds = self.inference_mixed_shapes(list(source))
                    elif im.is_cuda:
                        preds = self.inference_u8_device(im)
                    else:
                        preds = self.inference_from_host_u8(im, pipelined)
                    # a pipelined host batch lives on the output stream until its host copy has landed
                    host = self._host_copy(preds, self._out_stream if (pipelined and not mixed and not im.is_cuda) else None)
                if isinstance(source, (list, tuple)):
                    orig = list(source)
                else:                      # one [B,h,w,3] array / tensor: Results slice it on demand
                    orig = im if torch.is_tensor(source) else (source if source.ndim == 4 else source[None])
            else:
                if torch.is_tensor(source) and not source.is_cuda and source.dtype == torch.float32 and source.dim() == 4:
                    with profilers[0]:
                        im = check_tensor_source(source)
                    with profilers[1]:
                        preds = self.inference_from_host(im, pipelined)
                        if pipelined:
                            host = self._host_copy(preds, self._out_stream)
                else:
                    with profilers[0]:
                        im = self.preprocess(source)
                    with profilers[1]:
                        preds = self.inference(im)
                orig = im
        return (preds, im, orig, host, profilers)

    def _finish(self, ticket):
        """Wait for a queued batch's detections and wrap them (postprocess + callbacks of the reference's batch loop)."""
        preds, im, orig, host, profilers = ticket
        with torch.no_grad(), profilers[2]:
            self.results = self.postprocess(preds, im, orig, host=host)
        self.run_callbacks("on_predict_postprocess_end")
        n = max(len(self.results), 1)
        speed = {"preprocess": profilers[0].dt * 1e3 / n, "inference": profilers[1].dt * 1e3 / n,
                 "postprocess": profilers[2].dt * 1e3 / n}         # per-image averages of the batch, as the reference reports
        for r in self.results:
            r.speed = speed
        self.run_callbacks("on_predict_batch_end")
        return self.results


class Plan:
    """One recorded step of ``model.detect`` replayed from C (include/lpcyolo.h ``lpc_plan_*``): the layer loop, the concat
    planning and the side-stream forks run ONCE in Python while the library records every launch; ``run()`` then re-issues the
    launch sequence (as a CUDA graph captured inside the library) without any Python or torch code between the launches.
    The recording runs inside a private torch memory pool, so every activation of the step keeps its address for the lifetime
    of the plan; new inputs are written into ``self.x`` (the recorded input buffer), results appear in ``self.out``.

        plan = Plan(yolo.model, x_example)      # x: [B,3,H,W] float in [0,1] (or an NHWC activation), on the device
        plan.x.copy_(batch); plan.run(); dets = plan.out          # [B,max_det,6]
    """

    def __init__(self, model, x, max_det=300, clip=True):
        import ctypes as C
        from . import _lib
        if not x.is_cuda:
            raise F.LpcError("Plan: the example input must be on a CUDA device (no CPU fallback)")
        self._L = _lib.lib()
        self._h = C.c_void_p()
        dev = x.device
        with torch.no_grad(), torch.cuda.device(dev):
            model.detect(x, max_det, clip=clip)                  # builds the packed weights outside the plan's pool
            torch.cuda.synchronize(dev)
            self._pool = torch.cuda.MemPool()
            with torch.cuda.use_mem_pool(self._pool):
                self.x = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=dev)      # keeps an NHWC view's pixel pitch
                self.x.copy_(x)
                _lib.check(self._L.lpc_plan_begin(), "plan_begin")
                try:
                    self.out = model.detect(self.x, max_det, clip=clip)
                finally:
                    _lib.check(self._L.lpc_plan_end(C.byref(self._h)), "plan_end")
            torch.cuda.synchronize(dev)
        self.device = dev
        self.launches = int(self._L.lpc_plan_size(self._h))

    def run(self, graph=True, stream=None):
        """Re-issue the recorded launches on ``stream`` (default: torch's current stream of the plan's device)."""
        import ctypes as C
        from . import _lib
        with torch.cuda.device(self.device):
            st = C.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)
            fn = self._L.lpc_plan_run_graph if graph else self._L.lpc_plan_run
            _lib.check(fn(self._h, st), "plan_run")
        return self.out

    def __del__(self):
        try:
            if self._h:
                self._L.lpc_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


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
        self.predictor = None            # its captured CUDA graphs point at the old packed weights
        return self

    def add_callback(self, event, func):
        self.callbacks[event].append(func)

    def load_state_dict(self, sd, strict=True):
        self.predictor = None            # its captured CUDA graphs point at the old packed weights
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
        global _HALF_WARNED
        if "half" not in args and not args.get("fp32") and not _HALF_WARNED:
            # deliberate deviation from cfg/default.yaml:53 (half: False): documented in INTEGRATION.md, said once
            import warnings
            warnings.warn("lpc-yolo_b200: predict() computes in bf16 by default (the reference defaults to fp32, cfg/default.yaml:53); "
                          "pass half=False (or fp32=True) for the fp32 validation mode, half=True to silence this note", stacklevel=2)
            _HALF_WARNED = True
        if self.predictor is None or args != getattr(self, "_last_args", None):
            self.predictor = (predictor or YOLOv10DetectionPredictor)(overrides=args, _callbacks=self.callbacks)
            self.predictor.setup_model(self.model)
            self._last_args = dict(args)
        # engine/predictor.py:188-206: ``stream=True`` -> generator of Results, else the list.  ``batch`` (LoadPilAndNumpy's
        # batch size, data/build.py:157) splits a longer source into steps ONLY when it is passed: without it the whole source
        # is one batch (the reference's default of 1 would run image by image - same detections, 64 times the launches).
        n = self.predictor.source_len(source)
        if "batch" in args and n > int(args["batch"]):
            gen = self.predictor.stream_inference(source, int(args["batch"]))
            return gen if stream else list(gen)
        if stream:
            return iter(self.predictor(source))
        return self.predictor(source)

    __call__ = predict


YOLOv10 = YOLO
