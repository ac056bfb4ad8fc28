"""Mirror of ``yolov8/model/components.py`` (:69-209) of the reference."""
from ..modules import C2f, Bottleneck, Conv, DFL, MSBlock, SPPF, Upsample, yolo_params  # noqa: F401
