"""Helpers for the -m gpu tests: build the product decoder from the golden weights."""
import os

import numpy as np
import torch

from tests.helpers import pepper_weights, random_decoder_weights

_cache = {}


def pepper_decoder():
    if "pepper" not in _cache:
        from hortimapping_b200.decoder import Decoder
        W, b, codes = pepper_weights()
        dec = Decoder(W, b)
        from hortimapping_b200.decoder import calibration_rows
        # the recipe config_decoder uses (fewer rows under compute-sanitizer, where the fp32 calibration pass runs ~100x slower)
        dec.calibrate(calibration_rows(codes, 0.15, n=int(os.environ.get("HM_TEST_CAL_ROWS", "262144"))))
        _cache["pepper"] = dec
    return _cache["pepper"]


def random_decoder(seed=0):
    key = ("rand", seed)
    if key not in _cache:
        from hortimapping_b200.decoder import Decoder
        W, b = random_decoder_weights(seed)
        dec = Decoder(W, b)
        _cache[key] = (dec, W, b)
    return _cache[key]


def random_rows(n, seed=0, lat_std=0.1, box=0.1):
    g = np.random.default_rng(seed)
    z = g.normal(0, lat_std, (n, 32))
    x = (g.random((n, 3)) * 2 - 1) * box
    return np.concatenate([z, x], 1).astype(np.float32)
