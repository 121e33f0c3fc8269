"""ctypes binding of libhortimapping_b200_testing.so -- TEST-ONLY (tests/, scripts/probe_*.py).

The testing library is the product sources compiled with -DHM_TESTING plus csrc/testing/*.cu (build.py).  It exports the whole
product ABI AND the hm_debug_* hooks; a context created through it is independent of the product library's contexts.  Nothing
in the product path imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhortimapping_b200_testing.so")
DEBUG_EXPORTS = ["hm_debug_exp_pose", "hm_debug_huber_w2", "hm_debug_tc_selftest", "hm_debug_tc_wait_cycles", "hm_debug_tc_trace",
                 "hm_debug_tc_mma_rate", "hm_debug_tc_pair_probe", "hm_debug_tc_ingest", "hm_debug_tc_plan"]
_tlib = None


def lib() -> C.CDLL:
    global _tlib
    if _tlib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with hortimapping_b200.build.build_testing_library()")
        L = _lib.bind(C.CDLL(LIB_PATH))
        L.hm_debug_exp_pose.argtypes = [C.c_void_p, _lib.c_float_p, C.c_int, C.c_int, _lib.c_float_p]
        L.hm_debug_huber_w2.argtypes = [C.c_void_p, _lib.c_float_p, C.c_int, C.c_float, _lib.c_float_p]
        L.hm_debug_tc_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.hm_debug_tc_wait_cycles.argtypes = [C.c_void_p, C.c_void_p]
        L.hm_debug_tc_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hm_debug_tc_mma_rate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hm_debug_tc_pair_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.hm_debug_tc_ingest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.hm_debug_tc_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]      # host only: no context, no GPU
        _tlib = L
    return _tlib


def testing_decoder(weights, biases, device=None):
    """A Decoder whose context lives in the testing library (so the hm_debug_* hooks can be called on its handle)."""
    from .decoder import Decoder
    return Decoder(weights, biases, device=device, _library=lib())
