"""TEST INFRASTRUCTURE ONLY: numpy restatement of wild_completion/utils.py:23-109 (get_rays, get_render_data), pinned to
golden vectors produced by the unmodified reference (oracle/gen_golden_render_data.py -> tests/golden/render_data.npz).
Returns numpy arrays where the reference returns torch tensors on cfg['device']."""
import numpy as np


def get_rays(sampled_pixels, invK):
    """utils.py:23-38."""
    n = sampled_pixels.shape[0]
    u_hom = np.concatenate([sampled_pixels, np.ones((n, 1))], axis=-1)
    return (u_hom[:, None, :] * invK).sum(-1).astype(np.float32)


def get_render_data(submap_id, id_imgs, depth_imgs, cam_poses, img_size, invK, cfg, min_pix_count_match=400, max_bbx_size=300,
                    down_rate=1):
    """utils.py:39-109; the two np.random.choice calls (:77, :88) consume the global numpy RNG exactly as the reference does."""
    rd = {"frame_id": [], "T_wc": [], "rays_fg": [], "rays_bg": [], "depth_fg": [], "depth_bg": [], "pix_fg": [], "pix_bg": [], "count": 0}
    r = cfg["opt"]["render"]
    n_fg, n_bg, pad = r["n_fg_pix"], r["n_bg_pix"], r["n_bg_pad"]
    for img_id, id_img in id_imgs.items():
        depth = depth_imgs[img_id]
        mask = id_img == submap_id                                             # :50
        valid = mask & (depth > 0.)                                            # :51-52
        if int(valid.sum()) < min_pix_count_match:                             # :53-55
            continue
        mv, mu = np.where(valid)
        min_v, max_v = max(mv.min() - pad, 0), min(mv.max() + pad, img_size[0] - 1)      # :57-60
        min_u, max_u = max(mu.min() - pad, 0), min(mu.max() + pad, img_size[1] - 1)
        bh, bw = max_v - min_v + 1, max_u - min_u + 1
        if bh > max_bbx_size or bw > max_bbx_size:                             # :62-64
            print("Too large bbx, possibly wrong data association, skip this frame")
            continue
        hh = np.linspace(min_v, max_v, int(bh / down_rate)).astype(np.int32)   # :65-66
        ww = np.linspace(min_u, max_u, int(bw / down_rate)).astype(np.int32)
        vv = np.repeat(hh, ww.shape[0])                                        # row-major crop grid (:67-71)
        uu = np.tile(ww, hh.shape[0])
        bg = ~mask[vv, uu]                                                     # :72
        pix_bg = np.stack([uu[bg], vv[bg]], -1)
        d_bg = depth[vv[bg], uu[bg]]
        if pix_bg.shape[0] > n_bg:                                             # :75-79
            ind = np.random.choice(pix_bg.shape[0], n_bg, replace=False)
            pix_bg, d_bg = pix_bg[ind, :], d_bg[ind]
        fg = valid[vv, uu]                                                     # :81
        pix_fg = np.stack([uu[fg], vv[fg]], -1)
        d_fg = depth[vv[fg], uu[fg]]
        if pix_fg.shape[0] > n_fg:                                             # :86-90
            ind = np.random.choice(pix_fg.shape[0], n_fg, replace=False)
            pix_fg, d_fg = pix_fg[ind, :], d_fg[ind]
        rd["frame_id"].append(img_id)
        rd["rays_fg"].append(get_rays(pix_fg, invK))
        rd["rays_bg"].append(get_rays(pix_bg, invK))
        rd["depth_fg"].append(d_fg.astype(np.float32))
        rd["depth_bg"].append(d_bg.astype(np.float32))
        rd["T_wc"].append(np.asarray(cam_poses[img_id], np.float32))
        rd["pix_fg"].append(pix_fg)
        rd["pix_bg"].append(pix_bg)
        rd["count"] += 1
    return rd
