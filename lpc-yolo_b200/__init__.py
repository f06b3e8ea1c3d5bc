"""lpc-yolo_b200: B200-native (sm_100a) inference hot path of LPC-YOLO / YOLOv10.

The directory name carries a hyphen (it is the product name), so import it by string::

    import importlib
    lpc = importlib.import_module("lpc-yolo_b200")
    model = lpc.YOLO("yolov10n.yaml")
    results = model.predict(images)            # images: float tensor [B,3,H,W] in [0,1]
"""
from ._lib import LpcError, build, lib  # noqa: F401
from .engine import YOLO, YOLOv10, Boxes, Plan, Results, YOLOv10DetectionPredictor  # noqa: F401
from .nn.tasks import YOLOv10DetectionModel, parse_model, yaml_model_load  # noqa: F401

__version__ = "0.1.0"
