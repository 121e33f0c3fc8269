"""Multi-GPU plumbing: fruits are independent, so they are sharded over ranks (one process per GPU) with
no data-path collective; the only exchange is ONE all-gather of a fixed 49-float record per fruit
[latent 32 | T_ow 16 | iter_count 1] after the loop."""
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


def pack_records(latents: torch.Tensor, T_ow: torch.Tensor, iters: torch.Tensor) -> torch.Tensor:
    n = latents.shape[0]
    return torch.cat([latents.reshape(n, 32).float(), T_ow.reshape(n, 16).float(), iters.reshape(n, 1).float()], 1)


def unpack_records(rec: torch.Tensor):
    return rec[:, :32], rec[:, 32:48].reshape(-1, 4, 4), rec[:, 48].round().to(torch.int32)


def gather_records(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather the per-rank record blocks into the global [n_total, 49] table (fruit order).  Ranks may hold
    blocks that differ by one fruit, so blocks are padded to the largest one for the fixed-size collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, RECORD, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * mx, RECORD, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)
