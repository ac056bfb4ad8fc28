"""Mirror of ``yolov8/model/yolov8_head.py`` (:73-158) of the reference."""
from ..modules import Head  # noqa: F401
