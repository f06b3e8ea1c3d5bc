"""The reference's utils/ops.py functions on this path, each on a kernel of liblpcyolo: ``v10postprocess`` (:851-864,
``lpc_v10_postprocess``), ``xywh2xyxy`` (:402-421), ``clip_boxes`` (:305-324) and ``scale_boxes`` (:89-124) - the three
box helpers are modes of ``lpc_scale_boxes``.  The engine itself runs all of this inside the fused tail kernel
(``lpc_v10_decode_topk_scaled``); these serve code written against the reference's function API."""
import torch

from .. import _lib
from .. import functional as F
from ..functional import _fp, _stream


def v10postprocess(preds, max_det, nc=80):
    """preds [B,A,4+nc] -> boxes [B,K,4], scores [B,K], labels [B,K] (int64); scores descending, ties by
    ascending anchor*nc+class (the reference's tie order is unspecified)."""
    assert 4 + nc == preds.shape[-1]
    return F.v10_postprocess(preds, max_det, nc)


def _rows(boxes, what):
    if not (torch.is_tensor(boxes) and boxes.is_cuda):
        raise _lib.LpcError(f"{what} runs on CUDA tensors only (there is no CPU fallback)")
    if boxes.dtype != torch.float32 or boxes.stride(-1) != 1 or boxes.shape[-1] < 4:
        raise _lib.LpcError(f"{what}: fp32 rows with a unit-stride last dimension of >= 4 expected")
    flat = boxes.reshape(-1, boxes.shape[-1])
    if flat.data_ptr() != boxes.data_ptr() or (flat.shape[0] > 1 and flat.stride(0) < 4):
        raise _lib.LpcError(f"{what}: rows must be viewable as [n, row] without a copy (the update is in place)")
    return flat


def _launch(flat, xywh_in, pad, gain, clip, what):
    with torch.cuda.device(flat.device):
        _lib.check(_lib.lib().lpc_scale_boxes(_fp(flat), flat.shape[0], flat.stride(0) if flat.shape[0] > 1 else flat.shape[1], int(xywh_in),
                                              float(pad[0]), float(pad[1]), float(gain), float(clip[1]), float(clip[0]), _stream()), what)


def xywh2xyxy(x):
    assert x.shape[-1] == 4, f"input shape last dimension expected 4 but input shape is {x.shape}"
    y = x.detach().clone().contiguous()
    _launch(_rows(y, "xywh2xyxy"), True, (0, 0), 1.0, (0, 0), "xywh2xyxy")
    return y


def clip_boxes(boxes, shape):
    """In place, like the reference: x to [0, shape[1]], y to [0, shape[0]]."""
    _launch(_rows(boxes, "clip_boxes"), False, (0, 0), 1.0, shape, "clip_boxes")
    return boxes


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """In place: boxes in ``img1_shape`` (network input) coordinates -> ``img0_shape`` (original image), clipped."""
    if xywh:
        raise NotImplementedError("scale_boxes(xywh=True) is not on the predict path (predict.py:35 passes xyxy)")
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    _launch(_rows(boxes, "scale_boxes"), False, pad if padding else (0, 0), gain, img0_shape, "scale_boxes")
    return boxes
