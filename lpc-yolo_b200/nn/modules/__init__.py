"""The ultralytics module surface this path needs (same names as nn/modules/__init__.py exports)."""
from .block import C2f, C2fCIB, CIB, DFL, LPC, PSA, SCDown, SPCA, SPPF, Attention, Bottleneck, RepVGGDW, space_to_depth
from .conv import CBAM, ChannelAttention, Concat, Conv, SpatialAttention, Upsample, autopad
from .head import Detect, v10Detect

__all__ = ("Conv", "Concat", "CBAM", "ChannelAttention", "SpatialAttention", "Upsample", "autopad", "DFL", "SPPF",
           "C2f", "Bottleneck", "RepVGGDW", "CIB", "C2fCIB", "Attention", "PSA", "SCDown", "space_to_depth", "SPCA",
           "LPC", "Detect", "v10Detect")
