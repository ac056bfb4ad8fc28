import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """YMS_TEST_OPTIONS="name=value,..." applies library debug options (yolo_ms_b200._lib.set_debug_option) to this test process:
    how tests/test_gpu_ops.py::test_half_cta_mode_in_a_subprocess re-runs a selection of tests with an experiment switched on
    (the library itself never reads the environment)."""
    spec = os.environ.get("YMS_TEST_OPTIONS")
    if spec:
        from yolo_ms_b200 import _lib
        for item in spec.split(","):
            name, _, value = item.partition("=")
            _lib.set_debug_option(name.strip(), int(value))


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "yolov8"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
