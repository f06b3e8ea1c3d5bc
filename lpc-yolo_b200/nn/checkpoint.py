"""Real-weight ingestion: reference checkpoints -> this package's models.

Counterpart of the reference's ``torch_safe_load`` (nn/tasks.py:704-758), ``attempt_load_one_weight``
(:800-823), ``BaseModel.load`` (:226-241) and the Hugging Face ``from_pretrained`` route of
``models/yolov10/model.py:10`` (``PyTorchModelHubMixin`` -> ``config.json`` + ``model.safetensors``).

A reference ``.pt`` file is a pickle of live ``ultralytics`` module objects (``ckpt["model"]`` / ``ckpt["ema"]``
is a ``YOLOv10DetectionModel`` instance, saved ``.half()``, engine/trainer.py save_model).  The reference
un-pickles it by importing its own classes; this package does not contain them and must not execute them, so
the file is read with a *restricted* unpickler:

  * an explicit (module, name) allowlist resolves to real objects: the tensor rebuild functions of ``torch._utils``,
    typed storages / dtypes / ``torch.Size``, the layer CLASSES of ``torch.nn.modules.*``, ``collections.OrderedDict``,
    numpy's array reconstruction, ``pathlib`` paths and a safe ``builtins`` subset; lookups read
    ``sys.modules[module].__dict__[name]`` (no import, no attribute traversal) and dotted names are refused, so
    ``('torch.serialization', 'os.system')``-style protocol-4 gadgets do not resolve;
  * every other global - ``ultralytics.*`` first of all, but also un-listed ``torch`` functions such as ``torch.load``
    or ``torch.hub.load`` - resolves to an inert placeholder class that only records its constructor arguments and
    ``__dict__`` state; un-listed standard-library globals raise.  No foreign code runs.

The parameter / buffer tree of the placeholder graph is then walked exactly like ``nn.Module.state_dict``
(``_parameters``, persistent ``_buffers``, ``_modules``), the model is rebuilt from the checkpoint's own
``yaml`` dict with this package's parser, and the tensors are loaded with ``strict=True`` - the state_dict keys
of both implementations are identical by construction (tests/test_oracle_golden.py pins that).
"""
import json
import pickle
import struct
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch

_SAFE_BUILTINS = {"set", "frozenset", "list", "dict", "tuple", "slice", "range", "complex", "int", "float", "bool",
                  "str", "bytes", "bytearray", "object"}
_PASS_PREFIXES = ("torch", "collections", "numpy", "pathlib", "_codecs", "copyreg", "datetime", "functools")


class Placeholder:
    """Inert stand-in for a class (or function) this package does not provide."""

    _ref_module = _ref_name = "?"

    def __new__(cls, *args, **kwargs):
        self = object.__new__(cls)
        self.__dict__["_ctor_args"] = (args, kwargs)
        return self

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots)
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_state"] = state

    def __call__(self, *args, **kwargs):          # a placeholder used as a factory yields another inert object
        return _placeholder(self._ref_module, self._ref_name + "()")(*args, **kwargs)

    def __repr__(self):
        return f"<{self._ref_module}.{self._ref_name} placeholder>"


_placeholder_cache = {}


def _placeholder(module, name):
    key = (module, name)
    if key not in _placeholder_cache:
        _placeholder_cache[key] = type(name.rsplit(".", 1)[-1], (Placeholder,), {"_ref_module": module, "_ref_name": name})
    return _placeholder_cache[key]


_DILL_TYPES = {t.__name__: t for t in (dict, list, tuple, set, frozenset, int, float, bool, str, bytes, bytearray, slice,
                                        range, complex, object, type(None), OrderedDict)}


def _dill_load_type(name):
    """``dill._dill._load_type``: the reference's patched ``torch.save`` (utils/patches.py) pickles with dill when it
    is installed, and dill writes plain types through this helper.  Only harmless builtins resolve."""
    return _DILL_TYPES.get(name) or _placeholder("dill._dill._load_type", str(name))


# Explicit allowlist: (module, name) pairs that may resolve to a real object.  Everything else becomes an inert placeholder.
# Lookups go through ``sys.modules[module].__dict__[name]`` - no import, no attribute traversal: pickle protocol >= 4
# resolves DOTTED names by walking attributes (``('torch.serialization', 'os.system')``), which is how a module-prefix
# pass-list is bypassed; dotted names are refused outright.
_ALLOWED_FUNCS = {
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("_codecs", "encode"), ("copyreg", "_reconstructor"),
}
_ALLOWED_CLASSES = {
    ("collections", "OrderedDict"), ("torch", "Size"), ("torch", "device"), ("torch", "Tensor"),
    ("torch.nn.parameter", "Parameter"), ("numpy", "ndarray"), ("numpy", "dtype"),
    ("pathlib", "PosixPath"), ("pathlib", "PurePosixPath"), ("pathlib", "PureWindowsPath"), ("pathlib", "Path"),
    ("datetime", "datetime"), ("datetime", "date"), ("datetime", "timedelta"),
}
_PATH_REMAP = {("pathlib", "WindowsPath"): ("pathlib", "PureWindowsPath")}      # checkpoints written on Windows


def _resolve_allowed(module, name):
    """The real object for an allowlisted global, else None."""
    import sys
    module, name = _PATH_REMAP.get((module, name), (module, name))
    mod = sys.modules.get(module)
    if mod is None:
        return None
    obj = mod.__dict__.get(name)
    if obj is None:
        return None
    if (module, name) in _ALLOWED_FUNCS:
        return obj if callable(obj) else None
    if (module, name) in _ALLOWED_CLASSES:
        return obj if isinstance(obj, type) else None
    if module == "torch":
        # typed storages (torch.FloatStorage ...), dtype singletons (torch.float16 ...)
        if isinstance(obj, torch.dtype):
            return obj
        if isinstance(obj, type) and name.endswith("Storage") and getattr(obj, "__module__", "").split(".")[0] == "torch":
            return obj
        return None
    if module.startswith("torch.nn.modules."):
        # layer classes only: un-pickling creates them with __new__ + __dict__ update (or __init__ with plain arguments)
        if isinstance(obj, type) and issubclass(obj, torch.nn.Module) and getattr(obj, "__module__", "").startswith("torch.nn.modules."):
            return obj
        return None
    return None


class RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if not isinstance(module, str) or not isinstance(name, str):
            raise pickle.UnpicklingError("refusing a non-string global in a checkpoint")
        if (module, name) == ("dill._dill", "_load_type"):
            return _dill_load_type
        if module == "builtins":
            if name in _SAFE_BUILTINS:
                return getattr(__import__("builtins"), name)
            raise pickle.UnpicklingError(f"refusing builtins.{name} in a checkpoint")
        root = module.split(".", 1)[0]
        if "." in name:
            if root in _PASS_PREFIXES or root == "builtins":
                raise pickle.UnpicklingError(f"refusing dotted global {module}:{name} in a checkpoint")
            return _placeholder(module, name)
        obj = _resolve_allowed(module, name)
        if obj is not None:
            return obj
        if root in _PASS_PREFIXES and root not in ("torch", "numpy"):
            # a standard-library global that is not on the allowlist has no business in a weights file
            raise pickle.UnpicklingError(f"refusing {module}.{name} in a checkpoint")
        return _placeholder(module, name)      # ultralytics.*, un-listed torch / numpy names: inert


class _PickleModule:
    """The ``pickle_module`` duck type ``torch.load`` expects."""
    __name__ = "lpc_yolo_b200_restricted_pickle"
    Unpickler = RestrictedUnpickler
    UnpicklingError = pickle.UnpicklingError

    @staticmethod
    def load(f, **kw):
        return RestrictedUnpickler(f, **kw).load()


def check_suffix(file, suffix=".pt"):
    """utils/checks.py check_suffix as used by tasks.py:718."""
    if Path(str(file)).suffix.lower() != suffix:
        raise AssertionError(f"acceptable suffix is {suffix}, got '{file}'")


def torch_safe_load(weight):
    """tasks.py:704-758: ``(ckpt dict, file)``.  No download, no auto-install: a missing file raises."""
    check_suffix(weight, ".pt")
    file = Path(weight)
    if not file.is_file():
        raise FileNotFoundError(f"'{weight}' does not exist (there is no asset download in this package)")
    ckpt = torch.load(str(file), map_location="cpu", pickle_module=_PickleModule, weights_only=False)
    if not isinstance(ckpt, dict):           # tasks.py:750-756: a bare YOLO object saved with torch.save(model)
        ckpt = {"model": getattr(ckpt, "model", ckpt)}
    return ckpt, str(file)


def module_state_dict(m, prefix="", out=None):
    """nn.Module.state_dict over a graph of real torch modules and placeholders (float32 copies)."""
    out = OrderedDict() if out is None else out
    d = m.__dict__
    for name, p in (d.get("_parameters") or {}).items():
        if p is not None:
            out[prefix + name] = p.detach().float()
    skip = d.get("_non_persistent_buffers_set") or set()
    for name, b in (d.get("_buffers") or {}).items():
        if b is not None and name not in skip:
            out[prefix + name] = b.detach().float() if b.is_floating_point() else b.detach().clone()
    for name, sub in (d.get("_modules") or {}).items():
        if sub is not None:
            module_state_dict(sub, prefix + name + ".", out)
    return out


def intersect_dicts(da, db, exclude=()):
    """utils/torch_utils.py intersect_dicts: keys of ``da`` present in ``db`` with equal shapes."""
    return {k: v for k, v in da.items() if k in db and all(x not in k for x in exclude) and v.shape == db[k].shape}


def _names_of(obj, nc):
    names = getattr(obj, "names", None)
    if isinstance(names, (list, tuple)):
        names = dict(enumerate(names))
    if not isinstance(names, dict) or len(names) != nc:
        names = {i: f"{i}" for i in range(nc)}
    return {int(k): str(v) for k, v in names.items()}


def attempt_load_one_weight(weight, device=None, inplace=True, fuse=False):
    """tasks.py:800-823: ``(model, ckpt)``; the model is a YOLOv10DetectionModel of THIS package in eval mode,
    built from the checkpoint's ``yaml`` and loaded strictly from its tensors (EMA preferred, as the reference)."""
    from .tasks import YOLOv10DetectionModel
    ckpt, weight = torch_safe_load(weight)
    src = ckpt.get("ema") or ckpt["model"]
    cfg = getattr(src, "yaml", None)
    if not isinstance(cfg, dict):
        raise TypeError(f"'{weight}' holds no model yaml; cannot rebuild the layer table")
    head = cfg["head"][-1][2]
    if head != "v10Detect":
        raise NotImplementedError(f"'{weight}' is a '{head}' model; only YOLOv10 / LPC detection checkpoints are in scope")
    sd = module_state_dict(src)
    model = YOLOv10DetectionModel(dict(cfg), ch=cfg.get("ch", 3), verbose=False)
    model.load_state_dict(sd, strict=True)
    if hasattr(model, "invalidate"):
        model.invalidate()
    model.names = _names_of(src, cfg["nc"])
    args = ckpt.get("train_args") or {}
    model.args = dict(args) if isinstance(args, dict) else dict(getattr(args, "__dict__", {}))
    model.pt_path = weight
    model.task = "detect"
    model.eval()
    if device is not None:
        model.to(device)
    return model, ckpt


def load_into(model, weights, verbose=False):
    """BaseModel.load (tasks.py:226-241): transfer every tensor whose name and shape match; returns the count."""
    if isinstance(weights, (str, Path)):
        src = str(weights)
        if src.endswith(".safetensors"):
            csd = load_safetensors(src)
            csd = {(k[6:] if k.startswith("model.model.") else k): v for k, v in csd.items()}
        else:
            ckpt, _ = torch_safe_load(src)
            csd = module_state_dict(ckpt.get("ema") or ckpt["model"])
    elif isinstance(weights, dict) and "model" in weights and not torch.is_tensor(weights["model"]):
        csd = module_state_dict(weights.get("ema") or weights["model"])
    elif isinstance(weights, dict):
        csd = {k: v.float() if v.is_floating_point() else v for k, v in weights.items()}
    else:
        csd = weights.state_dict() if hasattr(weights, "state_dict") and not isinstance(weights, Placeholder) else module_state_dict(weights)
    own = model.state_dict()
    csd = intersect_dicts(csd, own)
    model.load_state_dict(csd, strict=False)
    if hasattr(model, "invalidate"):
        model.invalidate()
    if verbose:
        print(f"Transferred {len(csd)}/{len(own)} items from pretrained weights")
    return len(csd)


# ---- safetensors (the HF hub format of models/yolov10/model.py:10) -------------------------------------------
_ST_DTYPES = {"F64": np.float64, "F32": np.float32, "F16": np.float16, "I64": np.int64, "I32": np.int32, "I16": np.int16,
              "I8": np.int8, "U8": np.uint8, "BOOL": np.bool_}


def load_safetensors(path):
    """Minimal reader: u64 little-endian header length, JSON header {name: {dtype, shape, data_offsets}}, raw
    little-endian row-major data.  Floating tensors come back float32."""
    out = OrderedDict()
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        if n > 100 << 20:
            raise ValueError(f"'{path}': implausible safetensors header ({n} bytes)")
        header = json.loads(f.read(n).decode("utf-8"))
        base = f.read()
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        lo, hi = meta["data_offsets"]
        buf = base[lo:hi]
        if meta["dtype"] == "BF16":
            t = torch.frombuffer(bytearray(buf), dtype=torch.bfloat16).float()
        else:
            if meta["dtype"] not in _ST_DTYPES:
                raise ValueError(f"'{path}': unsupported dtype {meta['dtype']} for '{name}'")
            t = torch.from_numpy(np.frombuffer(buf, dtype=_ST_DTYPES[meta["dtype"]]).copy())
            if t.is_floating_point():
                t = t.float()
        out[name] = t.reshape(meta["shape"])
    return out


def save_safetensors(path, tensors, metadata=None):
    """Writer for the same format (used to round-trip packed checkpoints and by the tests)."""
    rev = {np.dtype(v).name: k for k, v in _ST_DTYPES.items()}
    header, blobs, off = {}, [], 0
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    for name, t in tensors.items():
        a = np.ascontiguousarray(t.detach().cpu().numpy())
        b = a.tobytes()
        header[name] = {"dtype": rev[a.dtype.name], "shape": list(t.shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b)
        off += len(b)
    h = json.dumps(header, separators=(",", ":")).encode("utf-8")
    h += b" " * (-len(h) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


def from_pretrained_dir(folder):
    """The local half of ``PyTorchModelHubMixin.from_pretrained``: ``config.json`` carries the ctor kwargs the
    reference pushes (models/yolov10/model.py:18-24: names, model = yaml file, task), ``model.safetensors`` the
    state_dict of the ``YOLOv10`` facade (keys ``model.model.<i>...``).  Returns ``(model, names, task)``."""
    from .tasks import YOLOv10DetectionModel
    folder = Path(folder)
    cfg = json.loads((folder / "config.json").read_text())
    yaml_file = Path(str(cfg.get("model", "yolov10n.yaml"))).with_suffix(".yaml").name
    if (folder / yaml_file).is_file():           # a custom table shipped next to the weights
        yaml_file = folder / yaml_file
    names = cfg.get("names")
    nc = len(names) if names else None
    model = YOLOv10DetectionModel(yaml_file, nc=nc, verbose=False)
    sd = load_safetensors(folder / "model.safetensors")
    sd = OrderedDict(((k[6:] if k.startswith("model.model.") else k), v) for k, v in sd.items())
    model.load_state_dict(sd, strict=True)
    if hasattr(model, "invalidate"):
        model.invalidate()
    if names:
        model.names = {int(k): str(v) for k, v in (names.items() if isinstance(names, dict) else enumerate(names))}
    model.eval()
    return model, model.names, cfg.get("task", "detect")
