"""Multi-GPU plumbing: fruits are independent, so they are sharded over ranks (one process per GPU) with
no data-path collective; the only exchange is ONE all-gather of a fixed 49-float record per fruit
[latent 32 | T_ow 16 | iter_count 1] after the loop; `optimize_sharded` appends the status word
(HM_STATUS_* bits) as a 50th float so that saturation / invalid-submap reports survive the gather."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

RECORD = 49


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of fruit indices of `rank` (blocks differ by at most one fruit)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(latents: torch.Tensor, T_ow: torch.Tensor, iters: torch.Tensor, status: torch.Tensor = None) -> torch.Tensor:
    n = latents.shape[0]
    cols = [latents.reshape(n, 32).float(), T_ow.reshape(n, 16).float(), iters.reshape(n, 1).float()]
    if status is not None:
        cols.append(status.reshape(n, 1).float())          # status words are < 2^24: exact in fp32
    return torch.cat(cols, 1)


def unpack_records(rec: torch.Tensor):
    out = (rec[:, :32], rec[:, 32:48].reshape(-1, 4, 4), rec[:, 48].round().to(torch.int32))
    if rec.shape[1] > RECORD:
        out = out + (rec[:, 49].round().to(torch.int32),)
    return out


def gather_records(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather the per-rank record blocks into the global [n_total, 49] table (fruit order).  Ranks may hold
    blocks that differ by one fruit, so blocks are padded to the largest one for the fixed-size collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    width = local.shape[1]
    pad = torch.zeros(mx, width, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * mx, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


def optimize_sharded(opt, latents: torch.Tensor, T_ow: torch.Tensor, points_w, render_datas=None, cube_radius=0.08, pose_known=False):
    """Sequence-level driver (host side): every rank holds the same list of fruits, optimises only its own
    contiguous block with ONE batched call (`Optimizer.shape_pose_joint_opt_batch`, or `shape_opt_deepsdf_batch` when
    `render_datas` is None) and ONE all-gather brings every fruit's (latent, T_ow, iter_count, status) to every rank, in fruit
    order.  Fruits are independent, so the result is bit-identical to a single-rank call on the whole list."""
    n = latents.shape[0]
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(n, rank, world)
    lat, T = latents[lo:hi].clone(), T_ow[lo:hi].clone()
    if hi > lo:
        if render_datas is None:
            lat, T, iters, status = opt.shape_opt_deepsdf_batch(lat, T, list(points_w[lo:hi]))
        else:
            import numpy as np
            cr = np.broadcast_to(np.asarray(cube_radius, np.float32), (n,))[lo:hi]
            pk = np.broadcast_to(np.asarray(pose_known, bool), (n,))[lo:hi]
            lat, T, iters, status = opt.shape_pose_joint_opt_batch(lat, T, list(render_datas[lo:hi]), list(points_w[lo:hi]), cr, pk)
    else:
        iters = torch.zeros(0, dtype=torch.int32, device=latents.device)
        status = torch.zeros(0, dtype=torch.int32, device=latents.device)
    rec = gather_records(pack_records(lat, T, iters, status), n)
    return unpack_records(rec)
