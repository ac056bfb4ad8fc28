"""yolo_ms_b200 -- B200-native (sm_100a) inference hot path of rafaelghiorzi/YOLO-MS:
backbone/neck/head forward, DFL head decode and batched class-aware NMS behind the reference's
``yolov8`` module API.  See DESIGN.md / INTEGRATION.md at the repository root.

Drop-in usage (replaces ``from yolov8.yolov8 import YOLOv8``):

    from yolo_ms_b200.yolov8 import YOLOv8
    from yolo_ms_b200.postprocess import postprocess
"""
from ._lib import YmsError, launch_count            # noqa: F401
from .modules import (C2f, Backbone, Bottleneck, Conv, DFL, Head, MSBlock, Neck, SPPF, Upsample,  # noqa: F401
                      YOLOv8, yolo_params)
from .postprocess import postprocess, postprocess_batched                       # noqa: F401

__all__ = ["YOLOv8", "Backbone", "Neck", "Head", "Conv", "Bottleneck", "C2f", "MSBlock", "SPPF", "Upsample", "DFL",
           "yolo_params", "postprocess", "postprocess_batched", "YmsError", "launch_count"]
