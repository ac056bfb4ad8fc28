"""Build libyms_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m yolo_ms_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libyms_b200.so")
SOURCES = ["capi.cu", "conv_gemm.cu", "conv3x3.cu", "stem_tc.cu", "glue.cu", "ms_fused.cu", "head_decode.cu", "nms.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newest_source() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "yms_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False, prof: bool = False) -> str:
    """prof=True builds libyms_b200_prof.so with -DYMS_PROF (per-role cycle counters in the conv kernels,
    used by scripts/role_prof.py through YMS_LIB=...); the product library never carries them."""
    lib = LIB.replace(".so", "_prof.so") if prof else LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _newest_source():
        return lib
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", "prof" if prof else "rel")
    os.makedirs(bdir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, *(["-DYMS_PROF"] if prof else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"== {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", lib, *objs, "-lcudart"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, prof="--prof" in sys.argv))
