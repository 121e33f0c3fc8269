"""world_size-2 gloo test of the N > 1 path's host logic: fruit sharding + the single all-gather of the
49-float result records (the device loop itself needs a GPU and is covered by -m gpu tests)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hortimapping_b200.shard import gather_records, pack_records, shard_range, unpack_records
    lo, hi = shard_range(n_total, rank, world)
    g = torch.Generator().manual_seed(1234)
    lat_all, T_all = torch.randn(n_total, 32, generator=g), torch.randn(n_total, 4, 4, generator=g)
    it_all = torch.arange(n_total, dtype=torch.int32) % 7
    # every rank "optimises" only its shard (stand-in: the identity), then ONE all-gather
    rec = gather_records(pack_records(lat_all[lo:hi], T_all[lo:hi], it_all[lo:hi]), n_total)
    lat, T, it = unpack_records(rec)
    ok = torch.equal(lat, lat_all) and torch.equal(T, T_all) and torch.equal(it, it_all)
    q.put((rank, bool(ok), tuple(rec.shape)))
    dist.destroy_process_group()


def test_two_rank_shard_and_allgather_matches_single_process():
    ctx = mp.get_context("spawn")
    for n_total in (8, 7):          # even and ragged split
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(60)
        assert all(ok for _, ok, _ in res), res
        assert all(shape == (n_total, 49) for _, _, shape in res)
