"""Golden vectors for wild_completion/utils.py:39-109 get_render_data, produced by the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY; run in the build container (needs /root/reference):  python oracle/gen_golden_render_data.py
Writes tests/golden/render_data.npz: synthetic submap-id / depth images (inputs) and, for several (fruit, settings) cases, every
array of the reference's render_data dict under a fixed np.random seed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
H, W = 240, 320


def make_scene(seed=0, n_frames=5):
    """Submap-id images with a few blobs (ids 1..5; 0 = background) that move between frames, depth with invalid (0) holes."""
    g = np.random.default_rng(seed)
    id_imgs, depth_imgs, poses = {}, {}, {}
    vv, uu = np.mgrid[0:H, 0:W]
    for k in range(n_frames):
        img = np.zeros((H, W), np.int32)
        depth = (0.4 + 0.0001 * ((vv * 3 + uu * 5 + k) % 64)).astype(np.float32)        # smooth pattern: compresses well
        blobs = [(1, 120 + 6 * k, 150 - 4 * k, 34, 28), (2, 60, 60 + 10 * k, 14, 20), (3, 200, 250, 30, 60 + 45 * (k == 2)),
                 (4, 30 + 3 * k, 290, 9, 9), (5, 100, 40, 3 + 4 * k, 45)]
        for (i, cv, cu, rv, ru) in blobs:
            m = ((vv - cv) / rv) ** 2 + ((uu - cu) / ru) ** 2 <= 1.0
            img[m] = i
            depth[m] = (0.3 + 0.01 * i + 0.0002 * ((vv[m] + 2 * uu[m]) % 32)).astype(np.float32)
        depth[(vv * 7 + uu * 13 + k * 5) % 17 == 0] = 0.0         # missing depth (about 6 % of the pixels)
        if k == 3:
            depth[img == 2] = 0.0                                   # fruit 2 has no valid depth in frame 3 -> skipped (:54)
        id_imgs[100 + k] = img
        depth_imgs[100 + k] = depth
        T = np.eye(4)
        T[:3, 3] = g.random(3)
        poses[100 + k] = T
    return id_imgs, depth_imgs, poses


CASES = [  # (name, submap_id, n_fg, n_bg, n_bg_pad, kwargs)
    ("f1", 1, 200, 200, 20, {}),
    ("f2_small", 2, 50, 400, 5, {"min_pix_count_match": 100}),
    ("f3_bigbbx", 3, 100, 100, 20, {"max_bbx_size": 150}),          # frame 2's box is too large -> skipped with a message (:61-63)
    ("f4_tiny", 4, 200, 200, 20, {}),                                # < 400 matching pixels everywhere -> no frames
    ("f5_down", 5, 64, 64, 12, {"min_pix_count_match": 50, "down_rate": 2}),
    ("f1_nosample", 1, 100000, 100000, 0, {}),                       # limits never reached -> no RNG call
]


def main():
    ref_shim.install()
    from wild_completion.utils import get_render_data
    id_imgs, depth_imgs, poses = make_scene()
    K = np.array([[300.0, 0, 160.5], [0, 301.5, 119.25], [0, 0, 1]])
    invK = np.linalg.inv(K)
    out = {"invK": invK, "img_size": np.array([H, W]), "frame_ids": np.array(sorted(id_imgs))}
    for fid in id_imgs:
        out[f"id_{fid}"] = id_imgs[fid]
        out[f"depth_{fid}"] = depth_imgs[fid]
        out[f"pose_{fid}"] = poses[fid]
    for name, sid, n_fg, n_bg, pad, kw in CASES:
        cfg = {"device": "cpu", "opt": {"render": {"n_fg_pix": n_fg, "n_bg_pix": n_bg, "n_bg_pad": pad}}}
        np.random.seed(1234)
        rd = get_render_data(sid, id_imgs, depth_imgs, poses, (H, W), invK, cfg, **kw)
        out[f"{name}_count"] = np.array(rd["count"])
        out[f"{name}_frame_id"] = np.array(rd["frame_id"], np.int64)
        for i in range(rd["count"]):
            for key in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg"):
                out[f"{name}_{key}_{i}"] = rd[key][i].numpy()
            out[f"{name}_pix_fg_{i}"] = np.asarray(rd["pix_fg"][i])
            out[f"{name}_pix_bg_{i}"] = np.asarray(rd["pix_bg"][i])
        print(name, "frames", rd["frame_id"], [tuple(r.shape) for r in rd["rays_fg"]], [tuple(r.shape) for r in rd["rays_bg"]])
    np.savez_compressed(os.path.join(GOLD, "render_data.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(GOLD, "render_data.npz")), "bytes")


if __name__ == "__main__":
    main()
