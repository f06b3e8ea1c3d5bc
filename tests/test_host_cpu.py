"""CPU-side checks of the host mirror: parser known answers (SURVEY.md 4.1 tier T1), state_dict key parity
with the oracle/reference inventory, packing algebra, and that the C-ABI library exports what the header declares."""
import ctypes
import hashlib
import importlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

MODELS = ["yolov10n", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x", "lpc"]


def _model(pkg, oracle, name):
    return pkg.YOLOv10DetectionModel(oracle.MODEL_FILES[name])


@pytest.mark.parametrize("name", MODELS)
def test_parser_matches_reference_inventory(pkg, oracle, name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    m = _model(pkg, oracle, name)
    sd = m.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    txt = "\n".join(f"{k}:{v}" for k, v in sorted(shapes.items()))
    assert hashlib.sha1(txt.encode()).hexdigest() == str(g["keys_sha1"])      # same keys, same shapes as the reference
    assert len(sd) == int(g["n_keys"])
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    assert m.save == list(g["save"])
    det = m.model[-1]
    assert [float(s) for s in det.stride] == list(g["strides"])
    assert [s[0].conv.in_channels for s in det.cv2] == list(g["head_ch"])
    # the oracle's synthetic weights load strictly
    layers, _, meta = oracle.load_layers(name)
    m.load_state_dict(oracle.synth_state_dict(oracle.param_shapes(layers), 0, meta["strides"], 80), strict=True)


def test_activation_quirk(pkg, oracle):
    """SURVEY.md finding 1: block modules use Mish, YAML-level Conv and the head use SiLU (yolov10n: 41 / 46 / 9)."""
    m = _model(pkg, oracle, "yolov10n")
    conv_mod = importlib.import_module("lpc-yolo_b200.nn.modules.conv")
    acts = [type(x.act).__name__ for x in m.modules() if isinstance(x, conv_mod.Conv)]
    rep = [x for x in m.modules() if type(x).__name__ == "RepVGGDW"]
    assert acts.count("Mish") == 46
    assert acts.count("SiLU") + len(rep) == 41
    assert acts.count("Identity") == 9 + 2 * len(rep) - 0 or acts.count("Identity") >= 9


def test_lpc_yaml_quirks(pkg, oracle):
    m = _model(pkg, oracle, "lpc")
    det = m.model[-1]
    assert det.f == [20, 23, 26] and [s[0].conv.in_channels for s in det.cv2] == [64, 192, 384]
    dest, live, fold, prefold, upfold = m._plan()
    assert upfold == {15: 17, 18: 20}      # both neck Upsample layers fold into the C2f behind their Concat
    assert fold == {2: 3, 5: 6, 8: 9, 11: 12}
    assert prefold == {1: 3, 4: 6}      # a stride-1 3x3 Conv whose ONLY consumer is a folded space_to_depth runs inside the C2f call (7, 10 also feed Concats)
    assert 27 not in live                       # dead layer is skipped
    assert m.yaml["scale"] == ""                # scale falls back to the first key
    assert dest[15] == (16, 0) and dest[10][0] == 16 and dest[22] == (23, 0) and dest[17][0] == 23


def test_bn_fold_algebra(pkg):
    pack = importlib.import_module("lpc-yolo_b200.pack")
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 16, 3, 1, 1, bias=False)
    bn = torch.nn.BatchNorm2d(16, eps=1e-3)
    bn.running_mean.uniform_(-1, 1); bn.running_var.uniform_(0.5, 2); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-1, 1)
    bn.eval()
    x = torch.randn(2, 8, 9, 9)
    w, b = pack.fold_bn(conv.weight, None, bn)
    ref = bn(conv(x))
    got = torch.nn.functional.conv2d(x.double(), w, b, 1, 1).float()
    assert (ref - got).abs().max() < 1e-5


def test_qkv_permutation(pkg):
    block = importlib.import_module("lpc-yolo_b200.nn.modules.block")
    a = block.Attention(288, num_heads=4)
    assert (a.key_dim, a.head_dim) == (36, 72)                       # yolov10m
    perm = a._qkv_perm()
    assert sorted(perm.tolist()) == list(range(288 + 2 * 36 * 4))
    per = 2 * 36 + 72
    assert perm[0] == 0 and perm[36] == per and perm[4 * 36] == 36 and perm[8 * 36] == 72 and perm[8 * 36 + 72] == per + 72


def test_abi_exports_every_declared_symbol(pkg):
    lib_mod = importlib.import_module("lpc-yolo_b200._lib")
    header = open(os.path.join(ROOT, "include", "lpcyolo.h")).read()
    declared = set(re.findall(r"\b(lpc_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib_mod.SIGNATURES), declared ^ set(lib_mod.SIGNATURES)
    h = ctypes.CDLL(lib_mod.build())
    for name in declared:
        assert hasattr(h, name), name
    assert lib_mod.lib().lpc_abi_version() == 1
    assert lib_mod.lib().lpc_conv2d_tc_kpad(16, 3) == 192 and lib_mod.lib().lpc_conv2d_tc_kpad(64, 1) == 64


def test_no_cpu_fallback(pkg, oracle):
    """The product path must fail loudly without a CUDA tensor."""
    m = _model(pkg, oracle, "yolov10n")
    with pytest.raises(pkg.LpcError):
        m(torch.rand(1, 3, 64, 64))
    ops = importlib.import_module("lpc-yolo_b200.utils.ops")
    with pytest.raises(pkg.LpcError):
        ops.v10postprocess(torch.rand(1, 400, 84), 300, 80)


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fn in os.walk(os.path.join(ROOT, "lpc-yolo_b200")):
        for f in fn:
            if f.endswith(".py"):
                s = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(import|from)\s+(oracle|lpc_oracle)", s, re.M):
                    bad.append(f)
    assert not bad


def test_chunk_plan_covers_the_batch():
    """engine._chunk_plan: chunk sizes of a host batch always add up to the batch; the first chunk is the small one."""
    import importlib
    eng = importlib.import_module("lpc-yolo_b200.engine")
    for B in (1, 2, 3, 8, 12, 16, 30, 32, 36, 40, 64, 100, 128, 256):
        plan = eng._chunk_plan(B)
        assert sum(plan) == B and all(v > 0 for v in plan)
        if len(plan) == 2:
            assert plan[0] <= plan[1]
    assert eng._chunk_plan(64) == [24, 40]


def test_stream_inference_queues_one_step_ahead(monkeypatch):
    """engine.stream_inference (host logic, no GPU): a source longer than one batch is cut into steps of ``batch`` images in
    order (ragged last step), step k+1 is QUEUED before the host waits for step k, results come back per image in source
    order, and the predictor callbacks fire as in the reference's loop (engine/predictor.py:208-283)."""
    import contextlib
    import importlib

    import numpy as np
    eng = importlib.import_module("lpc-yolo_b200.engine")
    log = []

    class Stub(eng.YOLOv10DetectionPredictor):
        def _enqueue(self, source, pipelined=False):
            assert pipelined
            log.append(("enqueue", len(source)))
            return list(source)

        def _finish(self, ticket):
            log.append(("finish", len(ticket)))
            return [int(t[0, 0, 0]) if hasattr(t, "shape") else t for t in ticket]

    monkeypatch.setattr(eng.torch.cuda, "device", lambda dev: contextlib.nullcontext())
    p = Stub()
    p.model, p.device = object(), "cpu"
    events = []
    p.add_callback("on_predict_start", lambda s: events.append("start"))
    p.add_callback("on_predict_end", lambda s: events.append("end"))
    src = np.arange(11, dtype=np.uint8).reshape(11, 1, 1, 1) * np.ones((1, 2, 2, 3), np.uint8)
    gen = p.stream_inference(src, 4)
    assert log == [] and events == []                      # a generator: nothing runs before the first next()
    assert next(gen) == 0
    assert log == [("enqueue", 4), ("enqueue", 4), ("finish", 4)]      # step 1 was queued before step 0 was waited for
    assert list(gen) == list(range(1, 11))
    assert log == [("enqueue", 4), ("enqueue", 4), ("finish", 4), ("enqueue", 3), ("finish", 4), ("finish", 3)]
    assert events == ["start", "end"]
    # a list source and a batch larger than the source
    del log[:]
    assert list(p.stream_inference([7, 8, 9], 8)) == [7, 8, 9] and log == [("enqueue", 3), ("finish", 3)]
    assert eng.YOLOv10DetectionPredictor.source_len(src) == 11 and eng.YOLOv10DetectionPredictor.source_len(src[0]) == 1
    assert eng.YOLOv10DetectionPredictor.source_len(torch.zeros(5, 3, 8, 8)) == 5
