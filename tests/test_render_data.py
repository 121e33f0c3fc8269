"""wild_completion/utils.py:39-109 get_render_data: the oracle restatement against the golden vectors of
the unmodified reference (CPU), and the device path against the same vectors (GPU) -- bit-exact: integer / byte work, and
the ray directions repeat the reference's fp64 arithmetic."""
import numpy as np
import pytest

from tests.helpers import load_npz

CASES = [("f1", 1, 200, 200, 20, {}), ("f2_small", 2, 50, 400, 5, {"min_pix_count_match": 100}),
         ("f3_bigbbx", 3, 100, 100, 20, {"max_bbx_size": 150}), ("f4_tiny", 4, 200, 200, 20, {}),
         ("f5_down", 5, 64, 64, 12, {"min_pix_count_match": 50, "down_rate": 2}), ("f1_nosample", 1, 100000, 100000, 0, {})]


def scene(g):
    fids = [int(x) for x in g["frame_ids"]]
    return ({f: g[f"id_{f}"] for f in fids}, {f: g[f"depth_{f}"] for f in fids}, {f: g[f"pose_{f}"] for f in fids},
            tuple(int(x) for x in g["img_size"]), g["invK"])


def check_case(g, name, rd, to_np):
    assert rd["count"] == int(g[f"{name}_count"])
    assert [int(x) for x in rd["frame_id"]] == [int(x) for x in g[f"{name}_frame_id"]]
    for i in range(rd["count"]):
        for key in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg"):
            got, ref = to_np(rd[key][i]), g[f"{name}_{key}_{i}"]
            assert got.dtype == np.float32 and got.shape == ref.shape, (name, key, i, got.shape, ref.shape)
            np.testing.assert_array_equal(got, ref, err_msg=f"{name} {key} frame {i}")
        np.testing.assert_array_equal(np.asarray(rd["pix_fg"][i]), g[f"{name}_pix_fg_{i}"])
        np.testing.assert_array_equal(np.asarray(rd["pix_bg"][i]), g[f"{name}_pix_bg_{i}"])


def test_oracle_get_render_data_matches_reference_golden(capsys):
    from oracle import render_data_oracle as RO
    g = load_npz("render_data")
    id_imgs, depth_imgs, poses, img_size, invK = scene(g)
    for name, sid, n_fg, n_bg, pad, kw in CASES:
        cfg = {"device": "cpu", "opt": {"render": {"n_fg_pix": n_fg, "n_bg_pix": n_bg, "n_bg_pad": pad}}}
        np.random.seed(1234)
        rd = RO.get_render_data(sid, id_imgs, depth_imgs, poses, img_size, invK, cfg, **kw)
        check_case(g, name, rd, np.asarray)
    assert "Too large bbx" in capsys.readouterr().out


@pytest.mark.gpu
def test_device_get_render_data_matches_reference_golden(capsys):
    import torch
    from hortimapping_b200.render_data import get_render_data, get_rays
    g = load_npz("render_data")
    id_imgs, depth_imgs, poses, img_size, invK = scene(g)
    for name, sid, n_fg, n_bg, pad, kw in CASES:
        cfg = {"device": "cuda", "opt": {"render": {"n_fg_pix": n_fg, "n_bg_pix": n_bg, "n_bg_pad": pad}}}
        np.random.seed(1234)
        rd = get_render_data(sid, id_imgs, depth_imgs, poses, img_size, invK, cfg, **kw)
        assert all(t.is_cuda and t.dtype == torch.float32 for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg") for t in rd[k])
        check_case(g, name, rd, lambda t: t.cpu().numpy())
    assert "Too large bbx" in capsys.readouterr().out
    # ids that do not occur / are out of range / other integer dtypes
    cfg = {"device": "cuda", "opt": {"render": {"n_fg_pix": 10, "n_bg_pix": 10, "n_bg_pad": 2}}}
    assert get_render_data(77, id_imgs, depth_imgs, poses, img_size, invK, cfg)["count"] == 0
    assert get_render_data(5000, id_imgs, depth_imgs, poses, img_size, invK, cfg)["count"] == 0
    u8 = {f: a.astype(np.uint8) for f, a in id_imgs.items()}
    np.random.seed(1234)
    rd8 = get_render_data(np.uint8(1), u8, depth_imgs, poses, img_size, invK,
                          {"device": "cuda", "opt": {"render": {"n_fg_pix": 200, "n_bg_pix": 200, "n_bg_pad": 20}}})
    check_case(g, "f1", rd8, lambda t: t.cpu().numpy())
    # get_rays alone (utils.py:23-38)
    from oracle import render_data_oracle as RO
    pix = np.stack([np.arange(0, 320, 7), np.arange(0, 320, 7) % 240], -1)
    np.testing.assert_array_equal(get_rays(pix, invK), RO.get_rays(pix, invK))


def test_compact_ids_host_logic():
    """Frames carry arbitrary integer ids (utils.py:50 compares with ==); the device table is indexed by compacted rows."""
    from hortimapping_b200.render_data import compact_ids, MAX_IDS
    img = np.array([[0, 5, 5], [70000, 5, -3]], np.int32)
    rows, row_of, exact = compact_ids(img)
    assert exact and row_of == {-3: 0, 0: 1, 5: 2, 70000: 3} and rows.dtype == np.int32
    np.testing.assert_array_equal(rows, [[1, 2, 2], [3, 2, 0]])
    for sid, r in row_of.items():
        np.testing.assert_array_equal(rows == r, img == sid)
    rows, row_of, exact = compact_ids(np.array([[0, 1.5], [2.0, 2.0]]))
    assert not exact and row_of[2] == 1 and rows[0, 1] == -1          # a non-integral pixel matches no integer id
    g = load_npz("render_data")
    fid = int(g["frame_ids"][0])
    rows, row_of, _ = compact_ids(g[f"id_{fid}"])
    np.testing.assert_array_equal(rows, g[f"id_{fid}"])                # ids 0..5 are already compact
    with pytest.raises(ValueError):
        compact_ids(np.arange(MAX_IDS + 1, dtype=np.int32).reshape(1, -1))
