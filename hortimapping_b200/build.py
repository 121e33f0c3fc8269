"""Builds the CUDA libraries in-tree with nvcc for sm_100a.

  libhortimapping_b200.so          the product: the C ABI of include/hortimapping_b200.h, nothing else
  libhortimapping_b200_testing.so  TEST-ONLY superset: the same sources compiled with -DHM_TESTING (timeline trace and wait-cycle
                                   counters inside the decoder kernel, hm_debug_* hooks that evaluate single device functions)
                                   plus csrc/testing/*.cu (bring-up probes and micro-benchmarks of the tcgen05 building blocks).
                                   Loaded only by tests/ and scripts/probe_*.py (hortimapping_b200/_testing.py).
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhortimapping_b200.so")
LIB_TESTING = os.path.join(HERE, "libhortimapping_b200_testing.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared", "-Xptxas", "-v"]


def sources(testing: bool = False):
    src = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if testing:
        src += sorted(glob.glob(os.path.join(CSRC, "testing", "*.cu")))
    return src


def _deps(testing: bool):
    return sources(testing) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def is_stale(lib: str = LIB, testing: bool = False) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in _deps(testing))


def _build(lib: str, testing: bool, force: bool, verbose: bool) -> str:
    if not force and not is_stale(lib, testing):
        return lib
    flags = FLAGS + (["-DHM_TESTING"] if testing else []) + os.environ.get("HM_EXTRA_NVCC_FLAGS", "").split()
    cmd = [NVCC] + flags + ["-o", lib] + sources(testing)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building {os.path.basename(lib)}")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(os.path.join(HERE, "build_ptxas_testing.log" if testing else "build_ptxas.log"), "w") as f:
        f.write(res.stderr)
    return lib


def build_library(force: bool = False, verbose: bool = False) -> str:
    return _build(LIB, False, force, verbose)


def build_testing_library(force: bool = False, verbose: bool = False) -> str:
    return _build(LIB_TESTING, True, force, verbose)


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_testing_library(force=True))
