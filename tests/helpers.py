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


def point_to_mesh_distance(points, verts, faces, k=12):
    """Distance from each point to a triangle mesh (exact closest point on the k nearest triangles by centroid; numpy, test use)."""
    from scipy.spatial import cKDTree
    P = np.asarray(points, np.float64)
    V, F = np.asarray(verts, np.float64), np.asarray(faces, np.int64)
    tri = V[F]                                                     # (T,3,3)
    _, idx = cKDTree(tri.mean(1)).query(P, k=min(k, len(F)))
    idx = idx.reshape(len(P), -1)
    best = np.full(len(P), np.inf)
    for j in range(idx.shape[1]):
        a, b, c = (tri[idx[:, j], i] for i in range(3))
        ab, ac, ap = b - a, c - a, P - a
        d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
        bp = P - b
        d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
        cp = P - c
        d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
        vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
        denom = np.where(np.abs(va + vb + vc) > 0, va + vb + vc, 1.0)
        v, w = vb / denom, vc / denom
        q = a + ab * v[:, None] + ac * w[:, None]                  # interior
        with np.errstate(divide="ignore", invalid="ignore"):
            t_ab = np.clip(d1 / np.where(d1 - d3 != 0, d1 - d3, 1), 0, 1)
            t_ac = np.clip(d2 / np.where(d2 - d6 != 0, d2 - d6, 1), 0, 1)
            t_bc = np.clip((d4 - d3) / np.where((d4 - d3) + (d5 - d6) != 0, (d4 - d3) + (d5 - d6), 1), 0, 1)
        cand = [q, a + ab * t_ab[:, None], a + ac * t_ac[:, None], b + (c - b) * t_bc[:, None], a, b, c]
        inside = (v >= 0) & (w >= 0) & (v + w <= 1)
        d = np.full(len(P), np.inf)
        for i, cc in enumerate(cand):
            dd = np.linalg.norm(P - cc, axis=1)
            if i == 0:
                dd = np.where(inside, dd, np.inf)
            d = np.minimum(d, dd)
        best = np.minimum(best, d)
    return best
