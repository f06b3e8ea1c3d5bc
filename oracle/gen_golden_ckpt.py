"""Generate tests/golden/yolov10n_tiny.pt (+ .npz, .yaml) with the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container:   python oracle/gen_golden_ckpt.py

The checkpoint is written the way the reference's trainer writes one (engine/trainer.py:479-506: a dict whose
``model`` entry is the pickled, ``.half()`` ``YOLOv10DetectionModel`` object plus ``train_args``), from a
width-0.25 / max-channels-512 / nc=20 variant of yolov10n.yaml (0.9 M parameters, so the fixture stays small), with
name-keyed synthetic weights and the reference's own BN calibration.  The .npz holds what the reference itself
answers for that file: ``YOLO(file).predict(x, conf=0)`` detections, names, a digest of every tensor.  The layer
table is also written as ``yolov10n_tiny.yaml`` so the tests can assemble a Hugging Face style folder
(``config.json`` + ``model.safetensors``, the layout ``PyTorchModelHubMixin.save_pretrained`` produces for
models/yolov10/model.py:10) with the ``safetensors`` library at test time instead of committing a second copy.
"""
import hashlib
import os
import sys
from copy import deepcopy

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import lpc_oracle as O  # noqa: E402
import ref_shim  # noqa: E402
from gen_golden import calibrate_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NAMES = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable", "dog",
         "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor"]
SIZE = 320


def tensor_digest(sd) -> str:
    h = hashlib.sha1()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().float().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    ref_shim.install()
    torch.set_num_threads(8)
    # The reference pins torch 2.0.1 (requirements.txt:1) where torch.load un-pickles objects by default; torch >= 2.6
    # flipped the default, so the reference's own torch_safe_load (nn/tasks.py:728) needs the old behaviour here.
    _load = torch.load
    torch.load = lambda *a, **k: _load(*a, **{"weights_only": False, **k})
    from ultralytics import YOLO
    from ultralytics.nn.tasks import YOLOv10DetectionModel, yaml_model_load

    ref_yaml = os.path.join(ref_shim.REF_ROOT, "ultralytics", "cfg", "models", "v10", "yolov10n.yaml")
    cfg = yaml_model_load(ref_yaml)
    cfg["scales"] = {"n": [0.33, 0.25, 512]}
    cfg["nc"] = len(NAMES)
    torch.manual_seed(0)
    model = YOLOv10DetectionModel(deepcopy(cfg), ch=3, verbose=False)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = O.synth_state_dict(shapes, 7, [float(s) for s in model.stride], len(NAMES))
    model.load_state_dict(sd, strict=True)
    calibrate_reference(model, O.calibration_batches())
    model.names = dict(enumerate(NAMES))
    model.args = {"task": "detect", "imgsz": 640, "data": "VOC.yaml", "model": "yolov10n.yaml"}

    pt = os.path.join(OUT, "yolov10n_tiny.pt")
    ckpt = {"epoch": -1, "best_fitness": None, "model": deepcopy(model).half(), "ema": None, "updates": None,
            "optimizer": None, "train_args": dict(model.args), "train_metrics": {"fitness": 0.0},
            "date": "2026-10-18T00:00:00", "version": "8.1.34"}
    torch.save(ckpt, pt)

    yolo = YOLO(pt)                                   # models/yolo/model.py:21-25 -> YOLOv10 -> Model._load
    loaded = {k: v.float() for k, v in yolo.model.state_dict().items()}      # before predict(): AutoBackend fuses in place
    n_params = sum(p.numel() for p in yolo.model.parameters())
    x = O.synth_input(2, SIZE, seed=5)
    res = yolo.predict(x, conf=0.0, verbose=False)
    dets = np.stack([r.boxes.data.cpu().numpy().astype(np.float32) for r in res])
    with torch.no_grad():
        y = yolo.model(x)["one2one"][0] if isinstance(yolo.model(x), dict) else yolo.model(x)[0]
    rec = {"dets": dets, "y": y.numpy().astype(np.float32)[:, :, ::5], "size": np.int64(SIZE), "seed": np.int64(5),
           "names": np.array(NAMES), "n_keys": np.int64(len(loaded)),
           "n_params": np.int64(n_params),
           "digest": np.array(tensor_digest(loaded)), "stride": yolo.model.stride.numpy().astype(np.float32)}
    np.savez_compressed(os.path.join(OUT, "yolov10n_tiny.npz"), **rec)

    with open(os.path.join(OUT, "yolov10n_tiny.yaml"), "w") as f:      # the same table as a YAML (HF config.json names it)
        import yaml
        yaml.safe_dump({k: v for k, v in cfg.items() if k not in ("yaml_file", "scale")}, f, sort_keys=False)
    print("wrote", pt, os.path.getsize(pt), "bytes; params", int(rec["n_params"]), "dets", dets.shape)


if __name__ == "__main__":
    main()
