"""Pin the CPU oracle against fixtures produced by the unmodified reference (oracle/gen_golden.py)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

MODELS = ["yolov10n", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x", "lpc"]
# SURVEY.md section 4.1 tier T1 known answers (measured from the reference)
PARAMS = dict(yolov10n=2775520, yolov10s=8128272, yolov10m=16576768, yolov10b=20574384,
              yolov10l=25888688, yolov10x=31808960, lpc=3968354)
NKEYS = dict(yolov10n=595, yolov10s=619, yolov10m=799, yolov10b=835, yolov10l=1027, yolov10x=1135, lpc=582)
HEAD_CH = dict(yolov10n=[64, 128, 256], yolov10s=[128, 256, 512], yolov10m=[192, 384, 576],
               yolov10b=[256, 512, 512], yolov10l=[256, 512, 512], yolov10x=[320, 640, 640], lpc=[64, 192, 384])


def _g(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


@pytest.mark.parametrize("name", MODELS)
def test_parser_known_answers(oracle, name):
    g = _g(name)
    layers, save, meta = oracle.load_layers(name)
    shapes = oracle.param_shapes(layers)
    sd = oracle.synth_state_dict(shapes, 0, meta["strides"], 80)
    n_params = sum(v.numel() for k, v in sd.items()
                   if "running_" not in k and "num_batches" not in k)
    assert n_params == PARAMS[name] == int(g["n_params"])
    assert len(shapes) == NKEYS[name] == int(g["n_keys"])
    txt = "\n".join(f"{k}:{tuple(v)}" for k, v in sorted(shapes.items()))
    assert hashlib.sha1(txt.encode()).hexdigest() == str(g["keys_sha1"])
    assert save == list(g["save"])
    assert save == ([7, 10, 14, 17, 20, 23, 26] if name == "lpc" else [4, 6, 10, 13, 16, 19, 22])
    assert layers[-1].args[1] == HEAD_CH[name] == list(g["head_ch"])
    assert meta["strides"] == [8.0, 16.0, 32.0] == list(g["strides"])


@pytest.mark.parametrize("name", ["yolov10n", "lpc", "yolov10m", "yolov10s"])
def test_forward_matches_reference(oracle, name):
    """Calibrated BN stats, decoded y and the final [300,6] detections vs the reference's outputs."""
    g = _g(name)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    m = oracle.build(name)
    for j, k in enumerate(g["bn_probe_keys"]):
        k = str(k)
        ref = torch.from_numpy(g[f"bn_probe_var_{j}"])
        assert (m.sd[k] - ref).abs().max() / ref.abs().max() < 1e-4
    x = oracle.synth_input(1, int(g["x_size"]))
    dets, aidx, y, raw = m.predict(x)
    ys = y[0, :, :: int(g["y_stride"])]
    ref = torch.from_numpy(g["y_small"])
    assert ys.shape == ref.shape
    assert (ys - ref).abs().max() / ref.abs().max() < 1e-4       # fp32 summation-order noise only
    if "raw_small_0" in g.files:
        for l in range(3):
            r = torch.from_numpy(g[f"raw_small_{l}"])
            assert (raw[l][0] - r).abs().max() / r.abs().max() < 1e-4
    # detections: same multiset of scores; boxes agree where ranks are unambiguous
    rd = torch.from_numpy(g["dets_small"])
    assert torch.allclose(dets[0, :, 4], rd[:, 4], rtol=1e-3, atol=1e-6)
    gap = (rd[:-1, 4] - rd[1:, 4]).abs()
    safe = torch.ones(300, dtype=torch.bool)
    safe[:-1] &= gap > 1e-4 * rd[:-1, 4]
    safe[1:] &= gap > 1e-4 * rd[:-1, 4]
    safe[-1] = False
    assert safe.sum() > 30
    assert torch.equal(dets[0, safe, 5], rd[safe, 5])
    # pre-clamp reference boxes vs ours (ours are clamped to the image, as scale_boxes does)
    rb = rd[safe, :4].clamp(0, int(g["x_size"]))
    assert (dets[0, safe, :4] - rb).abs().max() < 1e-2


def test_tail_on_reference_raw_maps(oracle):
    """Decode + v10postprocess restatement fed the REFERENCE's raw head maps: scores bit-exact except
    for exp rounding, labels and boxes equal."""
    for name in ("yolov10n", "lpc"):
        g = _g(name)
        raw = [torch.from_numpy(g[f"raw_small_{l}"])[None] for l in range(3)]
        y = oracle.decode(raw, [8.0, 16.0, 32.0], 80)
        ref_y = torch.from_numpy(g["y_small"])
        assert torch.equal(y[0, 4:], ref_y[4:])                        # sigmoid scores: bit-exact
        assert (y[0, :4] - ref_y[:4]).abs().max() < 5e-4               # boxes: op-order rounding (px)
        dets, _ = oracle.postprocess(ref_y[None], 300, 80)
        rd = torch.from_numpy(g["dets_small"])
        assert torch.equal(dets[0, :, 4], rd[:, 4])
        distinct = torch.ones(300, dtype=torch.bool)
        eq = rd[:-1, 4] == rd[1:, 4]
        distinct[:-1] &= ~eq
        distinct[1:] &= ~eq
        distinct[-1] = False   # the cut may fall inside a tie group
        assert torch.equal(dets[0, distinct, 5], rd[distinct, 5])
        assert torch.equal(dets[0, distinct, :4], rd[distinct, :4])


def test_predict_facade_640(oracle):
    """Oracle vs the reference's full YOLO(...).predict(tensor, conf=0) at 640x640 (BN folded by fuse())."""
    for name in ("yolov10n", "lpc"):
        g = _g(name)
        m = oracle.build(name)
        x = oracle.synth_input(1, 640)
        dets, _, y, _ = m.predict(x)
        ref_y = torch.from_numpy(g["y640_sample"])
        assert (y[0, :, ::97] - ref_y).abs().max() / ref_y.abs().max() < 1e-4
        rd = torch.from_numpy(g["dets_predict_640"])
        assert rd.shape == (300, 6)
        assert torch.allclose(dets[0, :, 4], rd[:, 4], rtol=2e-3, atol=1e-6)
        gap = (rd[:-1, 4] - rd[1:, 4]).abs()
        safe = torch.ones(300, dtype=torch.bool)
        safe[:-1] &= gap > 2e-4 * rd[:-1, 4]
        safe[1:] &= gap > 2e-4 * rd[:-1, 4]
        safe[-1] = False
        assert torch.equal(dets[0, safe, 5], rd[safe, 5])
        assert (dets[0, safe, :4] - rd[safe, :4]).abs().max() < 2e-2
