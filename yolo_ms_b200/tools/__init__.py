"""Callers on either side of the hot path: checkpoint ingestion and the batched
inference driver, mirroring the reference's ``yolov8/tools`` package."""
