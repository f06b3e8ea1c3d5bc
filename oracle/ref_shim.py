"""Make the unmodified reference importable in the build container (TEST INFRASTRUCTURE ONLY).

The reference imports ``matplotlib`` (utils/__init__.py:20) and ``timm`` (block.py:1331 ff.) at module
level although nothing on the YOLOv10 / LPC path uses them; neither is installed here.  This registers
inert stand-ins in ``sys.modules`` and puts ``/root/reference`` (or ``$LPC_REF``) on ``sys.path``.
Only ``oracle/gen_golden.py`` and the optional reference-parity tests use it; it never runs on the GPU box.
"""
import importlib.machinery
import os
import sys
import tempfile
import types

REF_ROOT = os.environ.get("LPC_REF", "/root/reference")


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ultralytics"))


def install():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "timm", "timm.layers", "timm.layers.create_act",
                 "timm.layers.norm", "timm.layers.create_conv2d", "timm.layers.helpers", "timm.layers.mlp"):
        if name not in sys.modules:
            m = _Inert(name)
            m.__path__ = []
            m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
            sys.modules[name] = m
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="yolo_cfg_"))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
