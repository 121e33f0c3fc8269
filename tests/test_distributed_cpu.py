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


class _StubOptimizer:
    """Stands in for hortimapping_b200.Optimizer on CPU: a deterministic per-fruit map, so that the sharded driver can be
    checked against a single-process call."""

    def shape_opt_deepsdf_batch(self, lat, T, pts):
        it = torch.tensor([p.shape[0] % 5 for p in pts], dtype=torch.int32)
        for i, p in enumerate(pts):
            lat[i] += float(np.asarray(p).sum())
            T[i] = T[i] * 2.0
        return lat, T, it, (it * 8 + 0x40 * (it % 2)).to(torch.int32)      # a recognisable status word per fruit

    def shape_pose_joint_opt_batch(self, lat, T, rds, pts, cr, pk):
        lat, T, it, st = self.shape_opt_deepsdf_batch(lat, T, pts)
        lat += torch.tensor(np.asarray(cr, np.float32)).reshape(-1, 1) + torch.tensor(np.asarray(pk, np.float32)).reshape(-1, 1)
        return lat, T, it + torch.tensor([len(r["T_wc"]) for r in rds], dtype=torch.int32), st


def _driver_worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hortimapping_b200.shard import optimize_sharded
    g = np.random.default_rng(5)
    lat0, T0 = torch.from_numpy(g.standard_normal((n_total, 32)).astype(np.float32)), torch.eye(4).repeat(n_total, 1, 1)
    pts = [g.standard_normal((3 + i, 3)).astype(np.float32) for i in range(n_total)]
    rds = [{"T_wc": [None] * (i % 3)} for i in range(n_total)]
    out = []
    for rd in (None, rds):
        lat, T, it, st = optimize_sharded(_StubOptimizer(), lat0, T0, pts, rd, cube_radius=0.08, pose_known=np.arange(n_total) % 2 == 0)
        out.append((lat.clone(), T.clone(), it.clone(), st.clone()))
    dist.destroy_process_group()
    # single-process result of the same call
    ref = []
    for rd in (None, rds):
        lat, T, it, st = optimize_sharded(_StubOptimizer(), lat0, T0, pts, rd, cube_radius=0.08, pose_known=np.arange(n_total) % 2 == 0)
        ref.append((lat, T, it, st))
    ok = all(torch.equal(a, b) for o, r in zip(out, ref) for a, b in zip(o, r))
    q.put((rank, bool(ok)))


def test_two_rank_sharded_driver_equals_single_process():
    ctx = mp.get_context("spawn")
    for n_total in (6, 5, 1):       # even, ragged, fewer fruits than ranks
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_driver_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(60)
        assert all(ok for _, ok in res), (n_total, res)
