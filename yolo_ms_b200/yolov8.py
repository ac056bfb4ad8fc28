"""Mirror of the reference module path ``yolov8/yolov8.py`` (class YOLOv8, :8-31)."""
from .modules import YOLOv8  # noqa: F401
