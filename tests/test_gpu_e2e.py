"""GPU parity tests, tiers T4-T6: whole network vs the CPU oracle and the reference-generated fixtures.

  T4 raw head / decoded y:  fp32 mode  max|d|/max|ref| <= 1e-5 ... measured against an fp64 run of the oracle
                            (the reference itself sits 3e-6 from fp64); we allow 2e-5 on
                            the deepest models where fp32 summation order alone moves the reference by 7e-5.
                            bf16 mode  l2-rel <= 2e-2, max-normalised <= 0.1 (reference's own bf16 drift: 9.7e-3 / 7.1e-2)
  T5 detections [B,300,6]:  fp32 mode: >= 99 % matched (class equal, box within 1e-2 px, score within 1e-5);
                            bf16: match rate reported, >= 60 % required at score gaps above bf16 resolution
  T6 API:                   YOLO(yaml).predict(tensor) returns Results with boxes.data [n,6], speed keys, callbacks.
"""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
ALL = ["yolov10n", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x", "lpc"]
_cache = {}


def _pair(pkg, oracle, name):
    """(oracle model with calibrated BN, product model holding the same state_dict)."""
    if name not in _cache:
        torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
        om = oracle.build(name)
        pm = pkg.YOLOv10DetectionModel(oracle.MODEL_FILES[name])
        pm.load_state_dict(om.sd, strict=True)
        _cache[name] = (om, pm.cuda().eval())
    return _cache[name]


def _norm_err(got, ref):
    return ((got - ref).abs().max() / ref.abs().max()).item(), ((got - ref).norm() / ref.norm()).item()


@pytest.mark.parametrize("name", ALL)
def test_fp32_mode_raw_and_y(pkg, oracle, name):
    om, pm = _pair(pkg, oracle, name)
    pm.compute_dtype = torch.float32
    x = oracle.synth_input(2, 160)
    with torch.no_grad():
        out = pm(x.cuda())["one2one"]
    y, raw = out[0].cpu(), [r.float().cpu() for r in out[1]]
    om64 = oracle.OracleModel(om.name, om.layers, om.save, om.meta, om.sd).to(torch.float64)
    y64, raw64 = om64.forward(x.double())
    tol = 1e-5 if name in ("yolov10n", "yolov10s", "lpc") else 2e-5
    for l in range(3):
        e, _ = _norm_err(raw[l].double(), raw64[l])
        assert e < tol, f"{name} raw level {l}: {e:.2e}"
    e, _ = _norm_err(y.double(), y64)
    assert e < tol, f"{name} y: {e:.2e}"
    # and against the reference's own output stored in the fixture (first image, same seed)
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    ref = torch.from_numpy(g["y_small"])
    x1 = oracle.synth_input(1, 160)
    with torch.no_grad():
        y1 = pm(x1.cuda())["one2one"][0].cpu()
    e, _ = _norm_err(y1[0, :, :: int(g["y_stride"])], ref)
    assert e < 1e-4, f"{name} vs reference fixture: {e:.2e}"


@pytest.mark.parametrize("name", ALL)
def test_bf16_mode_raw(pkg, oracle, name):
    om, pm = _pair(pkg, oracle, name)
    pm.compute_dtype = torch.bfloat16
    x = oracle.synth_input(2, 320)
    with torch.no_grad():
        out = pm(x.cuda())["one2one"]
    y, raw = out[0].cpu(), [r.float().cpu() for r in out[1]]
    oy, oraw = om.forward(x)
    worst = (0.0, 0.0)
    for l in range(3):
        e, l2 = _norm_err(raw[l], oraw[l])
        worst = (max(worst[0], e), max(worst[1], l2))
    print(f"{name}: bf16 raw head error max-normalised {worst[0]:.3e}, l2-rel {worst[1]:.3e}")
    assert worst[0] < 0.1 and worst[1] < 2e-2


def _match_rate(dets, odets, box_tol, score_tol):
    ok = 0
    for b in range(dets.shape[0]):
        used = set()
        for r in range(dets.shape[1]):
            d = dets[b, r]
            cand = ((odets[b, :, 5] == d[5]) & ((odets[b, :, :4] - d[:4]).abs().max(1).values < box_tol)
                    & ((odets[b, :, 4] - d[4]).abs() < score_tol)).nonzero().flatten().tolist()
            cand = [c for c in cand if c not in used]
            if cand:
                used.add(cand[0])
                ok += 1
    return ok / (dets.shape[0] * dets.shape[1])


@pytest.mark.parametrize("name", ["yolov10n", "lpc", "yolov10m"])
def test_detections_fp32(pkg, oracle, name):
    om, pm = _pair(pkg, oracle, name)
    pm.compute_dtype = torch.float32
    x = oracle.synth_input(2, 320)
    with torch.no_grad():
        dets = pm.detect(x.cuda(), 300).cpu()
    odets, _, _, _ = om.predict(x)
    rate = _match_rate(dets, odets, 1e-2, 1e-5)
    print(f"{name}: fp32 detection match rate {rate:.4f}")
    assert rate >= 0.99


@pytest.mark.parametrize("name", ["yolov10n", "lpc"])
def test_detections_bf16_and_reference_640(pkg, oracle, name):
    om, pm = _pair(pkg, oracle, name)
    x = oracle.synth_input(1, 640)
    pm.compute_dtype = torch.float32
    with torch.no_grad():
        d32 = pm.detect(x.cuda(), 300).cpu()
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    rd = torch.from_numpy(g["dets_predict_640"])[None]       # the reference's YOLO(...).predict(x, conf=0) output
    rate = _match_rate(d32, rd, 2e-2, 2e-5)
    print(f"{name}: fp32 vs reference predict() match rate {rate:.4f}")
    assert rate >= 0.97
    pm.compute_dtype = torch.bfloat16
    with torch.no_grad():
        d16 = pm.detect(x.cuda(), 300).cpu()
    rate16 = _match_rate(d16, rd, 4.0, 0.05 * rd[..., 4].max().item())
    print(f"{name}: bf16 vs reference predict() match rate (4 px, 5% score) {rate16:.4f}")
    assert rate16 >= 0.6


def test_predict_api(pkg, oracle):
    om, _ = _pair(pkg, oracle, "yolov10n")
    yolo = pkg.YOLO("yolov10n.yaml")
    yolo.load_state_dict(om.sd)
    fired = []
    for ev in ("on_predict_start", "on_predict_batch_start", "on_predict_postprocess_end", "on_predict_batch_end", "on_predict_end"):
        yolo.add_callback(ev, lambda p, ev=ev: fired.append(ev))
    x = oracle.synth_input(2, 320)
    res = yolo.predict(x, conf=0.0, half=False)
    assert len(res) == 2 and res[0].boxes.data.shape == (300, 6) and res[0].boxes.data.is_cuda
    assert set(res[0].speed) == {"preprocess", "inference", "postprocess"}
    assert len(fired) == 5
    assert res[0].orig_img.shape == (320, 320, 3) and res[0].orig_img.dtype == np.uint8
    odets, _, _, _ = om.predict(x)
    assert _match_rate(torch.stack([r.boxes.data.cpu() for r in res]), odets, 1e-2, 1e-5) >= 0.99
    # default conf=0.25 keeps a prefix; class filter; bad kwargs / bad shapes raise like the reference
    res = yolo.predict(x, half=False)
    assert all(len(r) <= 300 and (r.boxes.conf > 0.25).all() for r in res)
    with pytest.raises(SyntaxError):
        yolo.predict(x, not_an_arg=1)
    with pytest.raises(ValueError):
        yolo.predict(torch.rand(1, 3, 100, 100))
    # CUDA-graph replay of the whole path gives identical detections
    m = yolo.model
    m.compute_dtype = torch.bfloat16
    xs = x.cuda()
    with torch.no_grad():
        eager = m.detect(xs, 300).clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                m.detect(xs, 300)
        torch.cuda.current_stream().wait_stream(s)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = m.detect(xs, 300)
        gr.replay()
        torch.cuda.synchronize()
    assert torch.equal(out, eager)
