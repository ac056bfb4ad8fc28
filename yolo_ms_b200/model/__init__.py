"""Mirrors of the reference's ``yolov8/model`` modules (same names, same classes)."""
