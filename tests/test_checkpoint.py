"""Real-weight ingestion: reference ``.pt`` checkpoints and Hugging Face folders.

The fixture ``tests/golden/yolov10n_tiny.pt`` was written by the UNMODIFIED reference the way its trainer writes
checkpoints (oracle/gen_golden_ckpt.py); ``yolov10n_tiny.npz`` holds what the reference's own
``YOLO(file).predict(x, conf=0)`` answered for it.  CPU tests: the file loads with no ``ultralytics`` importable, every
tensor arrives bit-exactly (sha1 over the reference's loaded state_dict), hostile pickles stay inert, the oracle run on
the ingested weights reproduces the reference's detections.  GPU test: ``YOLO(file).predict`` through the CUDA path.
"""
import hashlib
import importlib
import json
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN

PT = os.path.join(GOLDEN, "yolov10n_tiny.pt")
NPZ = os.path.join(GOLDEN, "yolov10n_tiny.npz")
YAML = os.path.join(GOLDEN, "yolov10n_tiny.yaml")


def _digest(sd):
    h = hashlib.sha1()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().float().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def ckpt_mod(pkg):
    return importlib.import_module("lpc-yolo_b200.nn.checkpoint")


def test_pt_loads_without_the_reference(pkg):
    g = np.load(NPZ)
    yolo = pkg.YOLO(PT)
    assert "ultralytics" not in sys.modules            # nothing of the reference was imported to read its pickle
    m = yolo.model
    assert isinstance(m, pkg.YOLOv10DetectionModel)
    sd = m.state_dict()
    assert len(sd) == int(g["n_keys"]) and sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    assert _digest(sd) == str(g["digest"])             # every tensor bit-identical to what the reference loaded
    assert [yolo.names[i] for i in range(20)] == list(g["names"])
    assert m.stride.tolist() == g["stride"].tolist()
    assert m.args["task"] == "detect" and m.pt_path == PT and yolo.ckpt["version"] == "8.1.34"
    assert not m.training


def test_placeholders_keep_foreign_code_inert(ckpt_mod, tmp_path):
    """Anything that is not torch / collections / numpy resolves to an inert placeholder; dangerous builtins raise."""
    evil = b"cos\nsystem\n(S'touch /tmp/lpc_pwned'\ntR."           # os.system('touch ...') as a REDUCE
    obj = ckpt_mod.RestrictedUnpickler(__import__("io").BytesIO(evil)).load()
    assert isinstance(obj, ckpt_mod.Placeholder) and not os.path.exists("/tmp/lpc_pwned")
    with pytest.raises(pickle.UnpicklingError):
        ckpt_mod.RestrictedUnpickler(__import__("io").BytesIO(b"cbuiltins\neval\n(S'1+1'\ntR.")).load()
    with pytest.raises(AssertionError):
        ckpt_mod.torch_safe_load(str(tmp_path / "weights.onnx"))
    with pytest.raises(FileNotFoundError):
        ckpt_mod.torch_safe_load(str(tmp_path / "missing.pt"))


def _proto4_global(module, name, arg=None):
    """A protocol-4 pickle: STACK_GLOBAL (module, name) then REDUCE with one argument (or no argument)."""
    import io
    import pickletools  # noqa: F401  (documentation of the opcodes used below)
    def short_unicode(s):
        b = s.encode()
        return b"\x8c" + bytes([len(b)]) + b
    body = b"\x80\x04" + short_unicode(module) + short_unicode(name) + b"\x93"       # PROTO 4, STACK_GLOBAL
    if arg is None:
        body += b")R."                                                                   # EMPTY_TUPLE REDUCE STOP
    else:
        body += short_unicode(arg) + b"\x85R."                                           # TUPLE1 REDUCE STOP
    return io.BytesIO(body)


def test_dotted_and_unlisted_globals_do_not_resolve(ckpt_mod):
    """protocol >= 4 resolves dotted names by attribute traversal, so a module-prefix pass-list lets
    ('torch.serialization', 'os.getpid') or ('collections', '_sys.getrecursionlimit') execute.  Dotted names under the
    trusted roots are refused, un-listed torch functions (torch.load, torch.hub.load ...) are inert placeholders, un-listed
    standard-library globals raise."""
    U = ckpt_mod.RestrictedUnpickler
    for mod, name in (("torch.serialization", "os.getpid"), ("collections", "_sys.getrecursionlimit"),
                      ("torch", "serialization.os.getpid"), ("functools", "partial"), ("copyreg", "add_extension"),
                      ("builtins", "getattr"), ("builtins", "__import__")):
        with pytest.raises(pickle.UnpicklingError):
            U(_proto4_global(mod, name)).load()
    for mod, name in (("torch", "load"), ("torch.hub", "load"), ("torch.utils.collect_env", "run"), ("torch.serialization", "load"),
                      ("os", "system"), ("subprocess", "check_output"), ("numpy", "load"), ("torch.nn.modules.module", "register_module_forward_hook")):
        obj = U(_proto4_global(mod, name, "echo pwned > /tmp/lpc_pwned4")).load()
        assert isinstance(obj, ckpt_mod.Placeholder), (mod, name, obj)
    assert not os.path.exists("/tmp/lpc_pwned4")
    # the allowlist still resolves what real checkpoints need
    assert U(_proto4_global("collections", "OrderedDict")).load() == {}
    assert isinstance(U(_proto4_global("torch.nn.modules.activation", "SiLU")).load(), torch.nn.SiLU)
    assert U(_proto4_global("torch", "float16", None).__class__(b"\x80\x04\x8c\x05torch\x8c\x07float16\x93.")).load() is torch.float16


def test_oracle_on_ingested_weights_matches_reference_predict(pkg, oracle):
    """The checkpoint's tensors + the oracle's restatement == the reference's own predict() on that file."""
    g = np.load(NPZ)
    yolo = pkg.YOLO(PT)
    layers, save, meta = oracle.load_layers(YAML)
    om = oracle.OracleModel("tiny", layers, save, meta, yolo.model.state_dict())
    x = oracle.synth_input(2, int(g["size"]), seed=int(g["seed"]))
    dets, _, y, _ = om.predict(x)
    ref_y = torch.from_numpy(g["y"])
    assert ((y[:, :, ::5] - ref_y).abs().max() / ref_y.abs().max()).item() < 1e-4
    ref = torch.from_numpy(g["dets"])
    assert (dets[..., 5] == ref[..., 5]).float().mean().item() > 0.98
    same = dets[..., 5] == ref[..., 5]
    assert (dets[..., :4] - ref[..., :4]).abs().max(-1).values[same].max().item() < 2e-2
    assert ((dets[..., 4] - ref[..., 4]).abs() / ref[..., 4])[same].max().item() < 2e-4


def test_transfer_by_name_and_shape(pkg):
    """BaseModel.load (nn/tasks.py:226-241): same table -> everything transfers; other nc -> class convs are skipped."""
    full = pkg.YOLO(YAML)
    before = _digest(full.model.state_dict())
    full.load(PT)
    assert _digest(full.model.state_dict()) == str(np.load(NPZ)["digest"]) != before
    cfg = pkg.yaml_model_load(YAML)
    cfg["nc"] = 7
    other = pkg.YOLO(pkg.YOLOv10DetectionModel(cfg))
    ckpt_mod = importlib.import_module("lpc-yolo_b200.nn.checkpoint")
    n = ckpt_mod.load_into(other.model, PT)
    total = len(other.model.state_dict())
    assert 0 < total - n <= 12                         # only the 2x3 final class convs (weight+bias) differ in shape


def test_safetensors_reader_against_the_library(pkg, ckpt_mod, tmp_path):
    st = pytest.importorskip("safetensors.torch")
    yolo = pkg.YOLO(PT)
    sd = yolo.model.state_dict()
    # the facade's keys carry one more 'model.' level (models/yolov10/model.py:10: YOLOv10.model = DetectionModel)
    st.save_file({"model." + k: v.half().contiguous() if v.is_floating_point() else v.contiguous() for k, v in sd.items()},
                 str(tmp_path / "model.safetensors"), metadata={"format": "pt"})
    (tmp_path / "config.json").write_text(json.dumps({"names": yolo.names, "model": "yolov10n_tiny.yaml", "task": "detect"}))
    (tmp_path / "yolov10n_tiny.yaml").write_text(open(YAML).read())
    hub = pkg.YOLOv10.from_pretrained(tmp_path)
    assert _digest(hub.model.state_dict()) == _digest(sd)      # the checkpoint holds fp16 values, so half() is lossless
    assert hub.names == yolo.names
    # and our writer is readable by the library
    ckpt_mod.save_safetensors(str(tmp_path / "ours.safetensors"), {k: v for k, v in sd.items()}, {"format": "pt"})
    back = st.load_file(str(tmp_path / "ours.safetensors"))
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


@pytest.mark.gpu
def test_predict_from_checkpoint_on_gpu(pkg):
    from test_gpu_e2e import _match_rate
    import lpc_oracle as oracle
    g = np.load(NPZ)
    yolo = pkg.YOLO(PT)
    x = oracle.synth_input(2, int(g["size"]), seed=int(g["seed"]))
    res = yolo.predict(x, conf=0.0, half=False)
    dets = torch.stack([r.boxes.data.float().cpu() for r in res])
    ref = torch.from_numpy(g["dets"])
    rate = _match_rate(dets, ref, 2e-2, 2e-4)
    print(f"checkpoint predict (fp32 mode) vs reference predict(): match rate {rate:.4f}")
    assert rate >= 0.98
    assert res[0].names == yolo.names and res[0].names[14] == "person"
    res16 = yolo.predict(x, conf=0.0, half=True)
    d16 = torch.stack([r.boxes.data.float().cpu() for r in res16])
    print(f"checkpoint predict (bf16) same-set rate (4 px, 10 % score): {_match_rate(d16, ref, 4.0, 0.10):.3f}")
