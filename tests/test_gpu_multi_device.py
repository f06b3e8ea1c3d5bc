"""A model living on cuda:1 driven from a process whose current device is cuda:0 (the C library launches on the
CURRENT device's streams, and cudaFuncSetAttribute is a per-device setting).  Needs two GPUs; skipped otherwise."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_predict_on_a_non_current_device(pkg, oracle):
    synth = importlib.import_module("lpc-yolo_b200.utils.synth")
    torch.cuda.set_device(0)
    rng = np.random.default_rng(4)
    ims = [rng.integers(0, 256, (256, 256, 3), dtype=np.uint8) for _ in range(8)]
    x = oracle.synth_input(2, 320)
    outs = {}
    for dev in ("cuda:0", "cuda:1"):
        yolo = pkg.YOLO("yolov10-SPD-Conv-Tiny-CBAM-LPC.yaml")       # every kernel family incl. the large-shared-memory ones
        synth.init_synthetic(yolo.model)
        a = yolo.predict(ims, imgsz=256, conf=0.0, half=True, device=dev)          # host path: chunked graph replays
        b = yolo.predict(x, conf=0.0, half=True, device=dev)                       # tensor path
        assert torch.cuda.current_device() == 0
        assert all(r.boxes.data.device == torch.device(dev) for r in a + b)
        outs[dev] = (torch.stack([r.boxes.data.cpu() for r in a]), torch.stack([r.boxes.data.cpu() for r in b]))
    assert torch.equal(outs["cuda:0"][0], outs["cuda:1"][0]) and torch.equal(outs["cuda:0"][1], outs["cuda:1"][1])
