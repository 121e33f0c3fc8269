"""Shared loaders for the golden fixtures (tests/golden/, produced by oracle/gen_golden.py)."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLD, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def cfg_of(case, key="cfg_json"):
    return json.loads(bytes(case[key]).decode())


def pepper_weights():
    z = load_npz("sweetpepper_32")
    W = [z[f"W{l}"] for l in range(9)]
    b = [z[f"b{l}"] for l in range(9)]
    return W, b, z["latent_codes"]


def oracle_decoder(dtype=np.float32):
    from oracle import hm_oracle as O
    W, b, _ = pepper_weights()
    return O.DecoderOracle(W, b, (4,), dtype)


def random_decoder_weights(seed=0, alive=True):
    """A random decoder of the shipped architecture (8x512, latent_in=[4], lin3 -> 477) whose layers
    are all alive -- the shipped checkpoints have a dead layer 3 (pre-activation always negative),
    so they alone would not exercise layers 0-3."""
    rng = np.random.default_rng(seed)
    dims_in = [35, 512, 512, 512, 512, 512, 512, 512, 512]
    dims_out = [512, 512, 512, 477, 512, 512, 512, 512, 1]
    W, b = [], []
    for l in range(9):
        w = rng.normal(0, 1.0 / np.sqrt(dims_in[l]), (dims_out[l], dims_in[l])) * (1.6 if l < 8 else 0.3)
        if l == 0:
            w[:, 32:] *= 6.0          # xyz inputs are ~0.05 m, make them matter
        W.append(w.astype(np.float32))
        b.append((rng.normal(0, 0.05, dims_out[l]) + (0.02 if alive and l < 8 else 0)).astype(np.float32))
    return W, b


def render_data_of(case):
    n = int(case["n_frames"])
    rd = {k: [case[f"rd_{k}_{i}"] for i in range(n)] for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg")}
    return rd
