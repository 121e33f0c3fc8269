"""Host-side mirror of wild_completion/utils.py:39-109 `get_render_data` (and :23-38 `get_rays`) on top of the C ABI
(the step before the hot path).

Same signature and the same `render_data` dict.  A frame's submap-id and depth images are uploaded to the GPU once (cached per
image object, the host scripts call this function once per fruit with the same image dicts, test_wild_completion.py:133-226)
and ONE device pass builds the pixel count and bounding box of every id in the frame; per fruit only the crop is touched.
The random subsampling stays `np.random.choice` on the host, so a seeded run selects exactly the reference's pixels.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import check

MAX_IDS = 1024


def compact_ids(id_img: np.ndarray):
    """Map the ids of a frame onto 0 .. n-1 (device table rows).  Returns (int32 image of table rows, {id: row}, exact) where
    `exact` is False if the image holds non-integral values (they can never equal an integer submap id and get row -1).
    Ids are arbitrary integers in the reference (`submap_id_img == submap_id`, utils.py:50); the device table has MAX_IDS rows,
    so ids are compacted per frame instead of being used as indices."""
    ids = np.ascontiguousarray(id_img)
    as_int = ids.astype(np.int64, copy=False) if np.issubdtype(ids.dtype, np.integer) else np.rint(ids).astype(np.int64)
    integral = np.issubdtype(ids.dtype, np.integer) or bool(np.array_equal(as_int, ids))
    if not integral:
        mask = as_int == ids
    uniq, inv = np.unique(as_int, return_inverse=True)
    rows = inv.reshape(ids.shape).astype(np.int32)
    if not integral:
        rows = np.where(mask, rows, -1).astype(np.int32)
    if uniq.shape[0] > MAX_IDS:
        raise ValueError(f"a frame with {uniq.shape[0]} distinct ids exceeds the device table ({MAX_IDS} rows)")
    return rows, {int(v): i for i, v in enumerate(uniq)}, integral


class _Frame:
    """Device copies of one frame + the per-id (count, min_v, max_v, min_u, max_u) table."""

    def __init__(self, id_img: np.ndarray, depth_img: np.ndarray, dev: torch.device):
        L = _lib.lib()
        self.h, self.w = int(id_img.shape[0]), int(id_img.shape[1])
        rows, self.row_of, _ = compact_ids(id_img)
        self.id_dev = torch.from_numpy(rows).to(dev)
        self.depth_dev = torch.from_numpy(np.ascontiguousarray(depth_img, np.float32)).to(dev)
        n_ids = max(len(self.row_of), 1)
        table = torch.empty(n_ids, 5, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(L.hm_frame_id_bboxes(None, self.id_dev.data_ptr(), self.depth_dev.data_ptr(), self.h, self.w, n_ids, table.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "hm_frame_id_bboxes")
        self.table = table.cpu().numpy()


_frames: "dict[tuple, _Frame]" = {}


def _fingerprint(a: np.ndarray):
    """Cheap content guard of the frame cache: shape, dtype and a strided sample of the pixels (an image that is modified in
    place, or a new array that reuses a dead one's id(), must not be served from a stale device copy)."""
    a = np.asarray(a)
    flat = a.reshape(-1)
    step = max(1, flat.shape[0] // 4096)
    return (a.shape, str(a.dtype), hash(flat[::step].tobytes()))


def clear_frame_cache():
    """Drops every cached device copy of frames (they are otherwise dropped with their host arrays)."""
    _frames.clear()


def _frame_of(id_img, depth_img, dev) -> _Frame:
    key = (id(id_img), id(depth_img), str(dev))
    fp = (_fingerprint(id_img), _fingerprint(depth_img))
    fr = _frames.get(key)
    if fr is not None and fr.fingerprint != fp:
        fr = None
    if fr is None:
        fr = _Frame(id_img, depth_img, dev)
        fr.fingerprint = fp
        _frames[key] = fr
        for obj in (id_img, depth_img):                                      # drop the device copy with the host image
            try:
                weakref.finalize(obj, _frames.pop, key, None)
            except TypeError:
                pass
    return fr


def get_rays(sampled_pixels, invK, device="cuda"):
    """utils.py:23-38 on the device: (N,2) integer [u, v] pixels -> (N,3) float32 directions (numpy, like the reference)."""
    dev = torch.device(device)
    pix = torch.as_tensor(np.ascontiguousarray(sampled_pixels, np.int32)).to(dev)
    n = pix.shape[0]
    rays = torch.empty(n, 3, device=dev)
    d_in = torch.zeros(n, device=dev)
    d_out, p_out = torch.empty(n, device=dev), torch.empty(n, 2, dtype=torch.int32, device=dev)
    K = (C.c_double * 9)(*np.asarray(invK, np.float64).reshape(9))
    with torch.cuda.device(dev):
        check(_lib.lib().hm_gather_rays(None, pix.data_ptr(), d_in.data_ptr(), None, n, K, rays.data_ptr(), d_out.data_ptr(), p_out.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream), "hm_gather_rays")
    return rays.cpu().numpy()


def get_render_data(submap_id, id_imgs, depth_imgs, cam_poses, img_size, invK, cfg, min_pix_count_match=400, max_bbx_size=300,
                    down_rate=1):
    """utils.py:39-109.  Tensors of the result live on cfg['device'] (a CUDA device), float32, like the reference's."""
    L = _lib.lib()
    render_data = {"frame_id": [], "T_wc": [], "rays_fg": [], "rays_bg": [], "depth_fg": [], "depth_bg": [], "pix_fg": [], "pix_bg": [], "count": 0}
    cfg_render = cfg['opt']['render']
    fg_pix_count, bg_pix_count, bg_pad = cfg_render['n_fg_pix'], cfg_render['n_bg_pix'], cfg_render['n_bg_pad']
    dev = torch.device(cfg["device"] if str(cfg["device"]).startswith("cuda") else "cuda")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    K = (C.c_double * 9)(*np.asarray(invK, np.float64).reshape(9))
    sid = int(submap_id) if float(submap_id) == int(submap_id) else None
    for img_id, submap_id_img in id_imgs.items():
        fr = _frame_of(submap_id_img, depth_imgs[img_id], dev)
        row = fr.row_of.get(sid) if sid is not None else None
        if row is None:
            continue                                                         # no pixel carries this id (:53-55)
        count, min_mv, max_mv, min_mu, max_mu = (int(x) for x in fr.table[row])
        if count < min_pix_count_match:                                      # :53-55
            continue
        min_v = max(min_mv - bg_pad, 0)                                      # :57-60
        max_v = min(max_mv + bg_pad, img_size[0] - 1)
        min_u = max(min_mu - bg_pad, 0)
        max_u = min(max_mu + bg_pad, img_size[1] - 1)
        bbx_h, bbx_w = max_v - min_v + 1, max_u - min_u + 1
        if bbx_h > max_bbx_size or bbx_w > max_bbx_size:                     # :61-64
            print("Too large bbx, possibly wrong data association, skip this frame")
            continue
        hh = np.linspace(min_v, max_v, int(bbx_h / down_rate)).astype(np.int32)       # :65-66 (fp64 linspace, truncated)
        ww = np.linspace(min_u, max_u, int(bbx_w / down_rate)).astype(np.int32)
        crop_h, crop_w = hh.shape[0], ww.shape[0]
        if crop_h and crop_w and (int(hh.max()) >= fr.h or int(ww.max()) >= fr.w):
            # img_size from the config exceeds the uploaded image: the reference's numpy indexing raises here, so does this
            raise IndexError(f"crop row {int(hh.max())} / column {int(ww.max())} is outside the {fr.h} x {fr.w} image of frame {img_id} "
                             f"(img_size = {tuple(img_size)})")
        n = crop_h * crop_w
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            grid = torch.from_numpy(np.concatenate([hh, ww])).to(dev)
            pix = torch.empty(2, n, 2, dtype=torch.int32, device=dev)        # [bg | fg] candidates, [u, v]
            dep = torch.empty(2, n, dtype=torch.float32, device=dev)
            counts = torch.empty(2, dtype=torch.int32, device=dev)
            check(L.hm_crop_candidates(None, fr.id_dev.data_ptr(), fr.depth_dev.data_ptr(), fr.h, fr.w, row, grid.data_ptr(), crop_h,
                                       grid.data_ptr() + 4 * crop_h, crop_w, pix[0].data_ptr(), dep[0].data_ptr(), pix[1].data_ptr(),
                                       dep[1].data_ptr(), counts.data_ptr(), st), "hm_crop_candidates")
            n_bg, n_fg = (int(x) for x in counts.cpu().numpy())
            out = []
            for which, n_cand, limit in ((0, n_bg, bg_pix_count), (1, n_fg, fg_pix_count)):      # bg first, then fg (:75-79, :86-90)
                sel_ptr, k = None, n_cand
                if n_cand > limit:
                    sample_ind = np.random.choice(n_cand, limit, replace=False)
                    sel = torch.from_numpy(np.ascontiguousarray(sample_ind, np.int64)).to(dev)
                    sel_ptr, k = sel.data_ptr(), limit
                rays = torch.empty(k, 3, dtype=torch.float32, device=dev)
                d_sel = torch.empty(k, dtype=torch.float32, device=dev)
                p_sel = torch.empty(k, 2, dtype=torch.int32, device=dev)
                check(L.hm_gather_rays(None, pix[which].data_ptr(), dep[which].data_ptr(), sel_ptr, k, K, rays.data_ptr(), d_sel.data_ptr(),
                                       p_sel.data_ptr(), st), "hm_gather_rays")
                out.append((rays, d_sel, p_sel))
        (rays_bg, depth_bg, pix_bg), (rays_fg, depth_fg, pix_fg) = out
        render_data["frame_id"].append(img_id)
        render_data["rays_fg"].append(rays_fg)
        render_data["rays_bg"].append(rays_bg)
        render_data["depth_fg"].append(depth_fg)
        render_data["depth_bg"].append(depth_bg)
        render_data["T_wc"].append(torch.tensor(cam_poses[img_id], device=dev, dtype=torch.float32))
        render_data["pix_fg"].append(pix_fg.cpu().numpy())                   # just for vis (:103-105)
        render_data["pix_bg"].append(pix_bg.cpu().numpy())
        render_data["count"] += 1
    return render_data
