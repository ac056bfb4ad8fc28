"""CPU oracle for the YOLO-MS / YOLOv8 inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and the BASELINE legs of
``bench.py`` (``cpu_baseline`` / ``--impl reference``: the port on the host cores;
``gpu_library_baseline``: the same port's ATen ops run on the GPU through cuDNN +
``torchvision.ops.nms``, i.e. what the unmodified reference reaches on a CUDA device;
``verified``: the post-run self-check) may import it, and only as the checker or as
the thing timed *as a baseline* -- never as a fallback for the CUDA path
(``yolo_ms_b200`` raises if its CUDA library is missing).

Contents
--------
* ``yolov8_oracle``  functional fp32 restatement of the reference model forward
  (``/root/reference/yolov8/yolov8.py:23-31`` and the files it calls), plus the
  repo-local MS-Block definition (parity-unpinned: the reference has no MS-Block).
* ``postprocess``    restatement of the inline post-process of
  ``/root/reference/yolov8/tools/test.py:166-218`` and a numpy greedy NMS that
  restates ``torchvision.ops.nms`` (third-party, torchvision 0.26.0, not vendored
  by the reference; semantics pinned by executing the installed library).
* ``nms_ref.c``      the same greedy NMS in plain C for the 30k-box stress sizes.
* ``weights``        deterministic, manifest-driven state_dict generator.
* ``make_golden``    generates ``tests/golden/*`` by importing the REAL reference
  from ``/root/reference`` (only runnable in the build container).

Pinning status: YOLOv8 wiring, decode and post-process are pinned against the real
reference run in the build container (``tests/golden``, ``tests/test_oracle_*``).
The MS-Block variant is "parity unpinned" (no reference implementation exists).
"""
