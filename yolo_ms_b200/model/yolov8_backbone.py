"""Mirror of ``yolov8/model/yolov8_backbone.py`` (:34-74) of the reference."""
from ..modules import Backbone  # noqa: F401
