"""Mirror of ``yolov8/model/yolov8_neck.py`` (:55-94) of the reference."""
from ..modules import Neck  # noqa: F401
