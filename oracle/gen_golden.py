"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container (the reference is mounted at /root/reference; it does not exist on the GPU
box):   python oracle/gen_golden.py

For every v10 / LPC model YAML of the reference it
  * builds the reference model with the reference's own parser (nn/tasks.py:639 YOLOv10DetectionModel),
  * records parser known-answers (parameter total, state_dict keys+shapes digest, save list, head
    channels, strides),
  * loads the name-keyed synthetic weights of ``lpc_oracle.synth_state_dict`` (strict),
  * BN-calibrates with the reference's own BatchNorm2d in train mode (momentum=None),
  * runs eval-mode ``model(x)['one2one']`` on a seeded 160x160 input and the reference's
    ``ops.v10postprocess`` / ``ops.xywh2xyxy`` exactly as models/yolov10/predict.py:8-21 does,
  * for yolov10n and the LPC YAML additionally drives the full ``YOLO(...).predict`` facade at 640x640.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import lpc_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SMALL = 160


def keys_digest(shapes) -> str:
    txt = "\n".join(f"{k}:{tuple(v)}" for k, v in sorted(shapes.items()))
    return hashlib.sha1(txt.encode()).hexdigest()


def calibrate_reference(model, batches):
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None
    with torch.no_grad():
        for xb in batches:
            model(xb)
    model.eval()


def main():
    ref_shim.install()
    torch.set_num_threads(8)
    from ultralytics import YOLO
    from ultralytics.nn.tasks import YOLOv10DetectionModel
    from ultralytics.utils import ops

    os.makedirs(OUT, exist_ok=True)
    for name, fname in O.MODEL_FILES.items():
        ref_yaml = os.path.join(ref_shim.REF_ROOT, "ultralytics", "cfg", "models", "v10", fname)
        torch.manual_seed(0)
        model = YOLOv10DetectionModel(ref_yaml, ch=3, nc=80, verbose=False)
        ref_sd = model.state_dict()
        ref_shapes = {k: tuple(v.shape) for k, v in ref_sd.items()}
        layers, save, meta = O.load_layers(name)
        shapes = O.param_shapes(layers)
        assert {k: tuple(v) for k, v in shapes.items()} == ref_shapes, f"{name}: key/shape mismatch"
        assert list(shapes.keys()) == list(ref_sd.keys()), f"{name}: key order mismatch"
        assert save == model.save, (save, model.save)
        det = model.model[-1]
        assert [float(s) for s in det.stride] == [float(s) for s in meta["strides"]]
        n_params = sum(p.numel() for p in model.parameters())

        sd = O.synth_state_dict(shapes, 0, meta["strides"], 80)
        model.load_state_dict(sd, strict=True)
        calibrate_reference(model, O.calibration_batches())
        cal_sd = {k: v.clone() for k, v in model.state_dict().items()}

        x = O.synth_input(1, SMALL)
        with torch.no_grad():
            out = model(x)["one2one"]
        y, raw = out[0], out[1]
        preds = y.transpose(-1, -2)
        boxes, scores, labels = ops.v10postprocess(preds, 300, preds.shape[-1] - 4)
        dets = torch.cat([ops.xywh2xyxy(boxes), scores.unsqueeze(-1), labels.unsqueeze(-1)], -1)

        bn_keys = [k for k in cal_sd if k.endswith("running_var") and "one2one" not in k and ".cv2." not in k.split("model.")[1][:6]]
        probe = [bn_keys[0], bn_keys[len(bn_keys) // 2], [k for k in cal_sd if k.endswith("running_var") and "one2one_cv3" in k][-1]]
        rec = {
            "n_params": np.int64(n_params),
            "n_keys": np.int64(len(ref_sd)),
            "keys_sha1": np.array(keys_digest(ref_shapes)),
            "save": np.array(model.save, dtype=np.int64),
            "head_ch": np.array([m_.in_channels for m_ in [s[0].conv for s in det.cv2]], dtype=np.int64),
            "strides": np.array([float(s) for s in det.stride], dtype=np.float64),
            "x_size": np.int64(SMALL),
            "y_small": y[0].numpy().astype(np.float32) if name in ("yolov10n", "lpc", "yolov10m") else y[0, :, ::7].numpy().astype(np.float32),
            "y_stride": np.int64(1 if name in ("yolov10n", "lpc", "yolov10m") else 7),
            "dets_small": dets[0].numpy().astype(np.float32),
            "bn_probe_keys": np.array(probe),
        }
        for j, k in enumerate(probe):
            rec[f"bn_probe_var_{j}"] = cal_sd[k].numpy()
            rec[f"bn_probe_mean_{j}"] = cal_sd[k.replace("running_var", "running_mean")].numpy()
        if name in ("yolov10n", "lpc"):
            for l, r in enumerate(raw):
                rec[f"raw_small_{l}"] = r[0].numpy().astype(np.float32)
            # the full facade: YOLO(yaml).predict(tensor) (engine/model.py:385, predictor.py:208, predict.py:8)
            yolo = YOLO(ref_yaml)
            yolo.model.load_state_dict(cal_sd, strict=True)
            x640 = O.synth_input(1, 640)
            res = yolo.predict(x640, conf=0.0, verbose=False)
            rec["dets_predict_640"] = res[0].boxes.data.cpu().numpy().astype(np.float32)
            with torch.no_grad():
                y640 = model(x640)["one2one"][0]
            rec["y640_sample"] = y640[0, :, ::97].numpy().astype(np.float32)

        # cross-check the oracle right here so a broken fixture is never written silently
        om = O.build(name)
        for j, k in enumerate(probe):
            d = (om.sd[k] - cal_sd[k]).abs().max().item() / cal_sd[k].abs().max().item()
            assert d < 1e-4, (name, k, d)
        oy, _ = om.forward(x)
        err = (oy - y).abs().max().item() / y.abs().max().item()
        print(f"{name:9s} params={n_params} keys={len(ref_sd)} save={model.save} oracle-vs-ref max|d|/max|ref|={err:.2e}")
        assert err < 1e-4, (name, err)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)


if __name__ == "__main__":
    main()
