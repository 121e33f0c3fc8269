#!/usr/bin/env python
"""bench.py -- fruits/sec of the shape-completion inner loop on B200 (BASELINE.json metric).

Headline workload (BASELINE.json configs[1]): 64 synthetic fruit instances x 2048 observed surface points x
200 LM iterations per GPU, decoder-only path = `Optimizer.shape_opt_deepsdf`
(wild_completion/optimizer.py:306-429) batched over the fruits, all epsilon_* = 0 so no fruit exits
early ("200 Adam iters" in BASELINE.json is the reference's LM loop).

A step = one batched optimise call (64 fruits x 200 iterations).  `value` times K steps with the inputs
resident in HBM; `e2e` times the same K steps through the C-ABI host-buffer entry point
(hm_optimize_shape_host: pinned host inputs -> H2D -> loop -> D2H of latents/poses/iter counts).

The same JSON line carries a `joint` record: the FULL path north_star describes (`shape_pose_joint_opt`,
optimizer.py:28-302: render loss + recon loss + Sim(3) + latent), 32 fruits/GPU x (10 frames x 400 rays x 30
samples + 2048 points) x 200 LM iterations, with its own value / e2e / roofline (exact device-side row
counts) / cpu_baseline / parity.

At N = 1 rank 0 also runs the CPU baseline (`cpu_baseline`): the UNMODIFIED reference staged under
baseline/_ref (scripts/vendor_reference.py) on the host cores when it is there (kind "reference"), else the
numpy oracle port (kind "port"), in a subprocess, on fruit 0 of the same inputs; its result doubles as the in-run
parity check of the headline (`parity`: GPU latent after 200 iterations vs the CPU run's).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference|reference-cuda]
N > 1 runs under torchrun (one rank per GPU, fruits sharded, one NCCL all-gather of the 49-float result
records per step).  `--impl reference` times the reference's own CPU implementation alone (rank 0 only);
`--impl reference-cuda` the unmodified reference in PyTorch eager on cuda:0.
"""
from __future__ import annotations

import argparse
import copy
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

if "reference" in sys.argv or "--impl=reference" in sys.argv:
    # the CPU arm uses every host core, also under torchrun (which exports OMP_NUM_THREADS=1 to each rank):
    # the BLAS thread pools read these variables when numpy / torch are imported
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())
    # ... and it must not see a GPU at all: the reference creates tensors on "cuda" in places the shim's `.cuda()` identity does not
    # reach, so on a GPU box the CPU arm would end up with operands on two devices (this has to happen before torch is imported)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_FRUITS, N_PTS, N_ITERS = 64, 2048, 200
N_JOINT, N_JOINT_DISTINCT = 32, 8          # joint record: fruits per GPU (8 distinct synthetic fruits, each 4 times)
FLOP_PER_FWD_ROW = 3_671_040           # 2 x 1 835 520 MAC
FLOP_PER_JAC_ROW = 7_342_080           # forward + the same MAC count backward to the input
METRIC = "fruits/sec (200 iters, 2048 pts)"
PARITY_TOL = 1e-4                      # north_star: 1e-4 fp32 relative tolerance
REPLAY_STEPS = 5                       # LM iterations replayed one by one from the fp64 run's own states (the convergent phase)
OBJECTIVE_TOL = 1e-2                   # runs past convergence are compared by the objective they reached; the reference's own
                                       # fp32 / fp64 runs and its 30- / 200-iteration states differ by ~2e-3 in this measure

WILD_CFG = {
    "device": "cuda",
    "opt": {"scale_on": True,
            "lm": {"lm_on": True, "lm_eye": False, "lm_lambda_0": 0.1, "s_damp": 1e-3},
            "recon": {"n_pts": 2000, "cluster_dist_m": 0.01, "robust_th_m": 0.01},
            "render": {"n_fg_pix": 200, "n_bg_pix": 200, "n_bg_pad": 20, "n_frame": 10, "n_sample_on_ray": 30,
                       "log_sdf_occ": True, "occ_cutoff_m": 0.01, "occlusion_on": True, "robust_th_m": 0.05},
            "weight": {"w_recon": 1, "w_depth": 5e-2, "w_mask": 5e-4, "w_codereg": 5e-4},
            "converge": {"max_iter": N_ITERS, "epsilon_g": 0, "epsilon_c": 0, "epsilon_t": 0, "epsilon_r": 0, "epsilon_s": 0},
            "robust_iter": 5},
    "vis": {"log_on": False, "vis_on": False, "vis_pause_s": 0.0, "object_radius_max_m": 0.08, "mc_res_mm": 4.0},
}   # configs/wild_pepper.yaml of the reference with max_iter 200, epsilons 0, vis off


def load_weights():
    z = np.load(os.path.join(ROOT, "tests", "golden", "sweetpepper_32.npz"))
    return [z[f"W{l}"] for l in range(9)], [z[f"b{l}"] for l in range(9)], z["latent_codes"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        pw = sorted(float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit())
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw[len(pw) // 2:]) if pw else None}      # board power under load (the box runs at its 1000 W cap)


def product_sdf_jac(dec):
    import torch

    def sdf_jac(latent, pts):
        y, g = dec.sdf_jacobian(torch.from_numpy(np.asarray(latent, np.float32)), torch.from_numpy(np.asarray(pts, np.float32)))
        return y.reshape(-1).cpu().numpy(), g.reshape(-1, 35)[:, 32:].cpu().numpy()
    return sdf_jac


def make_inputs(dec, codes, seed: int, rank: int):
    """Synthetic fruits of the SURVEY 8d generator (points only), built with the product decoder."""
    from hortimapping_b200 import synth
    pts, T_ow = [], []
    if os.environ.get("HM_BENCH_RANDOM_POINTS"):       # profiling aid: skip the surface projection launches
        g = np.random.default_rng(seed + rank)
        p = ((g.random((N_FRUITS, N_PTS, 3)) * 2 - 1) * 0.045).astype(np.float32)
        return p, np.tile(np.eye(4, dtype=np.float32), (N_FRUITS, 1, 1)), np.tile(codes.mean(0).astype(np.float32), (N_FRUITS, 1))
    sj = product_sdf_jac(dec)
    for i in range(N_FRUITS):
        fr = synth.make_fruit(sj, codes, seed, rank * N_FRUITS + i, n_pts=N_PTS, with_rays=False)
        pts.append(fr.points_w)
        T_ow.append(np.linalg.inv(fr.T_wo_gt.astype(np.float64)).astype(np.float32))   # pose known for the DeepSDF baseline
    init_lat = np.tile(codes.mean(0).astype(np.float32), (N_FRUITS, 1))
    return np.stack(pts), np.stack(T_ow), init_lat


def make_joint_inputs(dec, codes, seed: int, rank: int):
    """SURVEY 8d generator with rays: 10 frames x (200 fg + 200 bg) rays, 2048 points, 20 % of the background rays carry
    a depth in front of the fruit (a synthetic leaf) so that the occlusion branch of loss.py:132-149 is exercised."""
    from hortimapping_b200 import synth
    sj = product_sdf_jac(dec)
    return [synth.make_fruit(sj, codes, seed, rank * N_JOINT_DISTINCT + i, n_pts=N_PTS, with_rays=True, leaf_fraction=0.2)
            for i in range(N_JOINT_DISTINCT)]


# ------------------------------------------------------------------------------------------------
# CPU / reference arms (never on the product path)
# ------------------------------------------------------------------------------------------------
def _torch_mm():
    import torch
    torch.set_num_threads(os.cpu_count())

    def mm(x, y):
        return torch.from_numpy(np.ascontiguousarray(x)).matmul(torch.from_numpy(np.ascontiguousarray(y))).numpy()
    return mm, torch.get_num_threads()


def cpu_arm(io: dict, n_shape_iters: int, n_joint_iters: int, want_states: bool, device: str = "cpu") -> dict:
    """Runs the reference's own implementation of the path on `io` (fruit 0 of the bench inputs, or synthetic points):
    the unmodified reference through oracle/ref_runner.py when it is staged, else the numpy oracle port with its GEMMs on
    torch's CPU BLAS (the library the reference itself runs on).  Returns timings and the resulting states."""
    from oracle import ref_runner
    out = {"device": device}
    cfg = copy.deepcopy(WILD_CFG)
    if ref_runner.reference_root() is not None:
        R = ref_runner.ReferenceRunner(device=device)
        out.update(kind="reference", cores=R.threads,
                   what=f"unmodified reference ({os.path.relpath(R.root, ROOT) if R.root.startswith(ROOT) else R.root}) via oracle/ref_shim.py, "
                        f"PyTorch {'CPU, ' + str(R.threads) + ' threads' if device == 'cpu' else 'eager on ' + device}")
        if n_shape_iters:
            R.shape_opt(cfg, io["init_lat"], io["T_ow"], io["points_w"], 2)       # warm-up (allocator, BLAS thread pool)
            lat, it, dt = R.shape_opt(cfg, io["init_lat"], io["T_ow"], io["points_w"], n_shape_iters)
            out.update(shape_s=dt, shape_iters=it, shape_latent=lat.tolist())
            if want_states:
                for k in (5, 30):
                    lat_k, _, _ = R.shape_opt(cfg, io["init_lat"], io["T_ow"], io["points_w"], k)
                    out[f"shape_latent{k}"] = lat_k.tolist()
        if n_joint_iters and "rd" in io:
            args = (io["j_init_lat"], io["j_init_T"], io["rd"], io["j_points_w"], 0.08, False)
            if want_states:
                ob = R.observed_joint_iteration(cfg, *args)
                out["joint_it0"] = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in ob.items()}
            lat, T, it, dt = R.joint_opt(cfg, *args, n_joint_iters)
            out.update(joint_s=dt, joint_iters=it)
        if want_states:
            probes = {k[len("probe_"):]: io[k] for k in io if k.startswith("probe_")}      # states of the B200 run, sent by the parent
            if n_shape_iters:
                probes.update(reference_30=np.asarray(out["shape_latent30"]), reference_200=np.asarray(out["shape_latent"]))
            out["fp64"] = fp64_truth(io, n_shape_iters, probes)
        return out
    if device != "cpu":
        raise RuntimeError("reference-cuda needs the unmodified reference under baseline/_ref (scripts/vendor_reference.py)")
    from oracle import hm_oracle as O
    mm, threads = _torch_mm()
    O.set_matmul(mm)
    W, b, _ = load_weights()
    dec = O.DecoderOracle(W, b, (4,), np.float32)
    out.update(kind="port", cores=threads, what=f"oracle/hm_oracle.py (numpy restatement, GEMMs on torch's CPU BLAS, {threads} threads)")
    if n_shape_iters:
        c = copy.deepcopy(cfg)
        c["opt"]["converge"]["max_iter"] = 2
        O.shape_opt_deepsdf(dec, c, io["init_lat"].copy(), io["T_ow"], io["points_w"])
        c["opt"]["converge"]["max_iter"] = n_shape_iters
        lat = io["init_lat"].copy()
        t0 = time.perf_counter()
        _, _, it = O.shape_opt_deepsdf(dec, c, lat, io["T_ow"], io["points_w"])
        out.update(shape_s=time.perf_counter() - t0, shape_iters=it, shape_latent=lat.tolist())
    if n_joint_iters and "rd" in io:
        c = copy.deepcopy(cfg)
        c["opt"]["converge"]["max_iter"] = n_joint_iters
        lat = io["j_init_lat"].copy()
        tr = O.OptTrace()
        t0 = time.perf_counter()
        _, _, it = O.shape_pose_joint_opt(dec, c, lat, io["j_init_T"], io["rd"], io["j_points_w"], 0.08, False, trace=tr)
        out.update(joint_s=time.perf_counter() - t0, joint_iters=it)
        if want_states:
            c1 = copy.deepcopy(cfg)
            c1["opt"]["converge"]["max_iter"] = 1
            t1 = O.OptTrace()
            l1 = io["j_init_lat"].copy()
            _, T1, _ = O.shape_pose_joint_opt(dec, c1, l1, io["j_init_T"], io["rd"], io["j_points_w"], 0.08, False, trace=t1)
            out["joint_it0"] = {"H": t1.H[0].tolist(), "b": t1.b[0].tolist(), "dx": t1.dx[0].tolist(), "rows_fwd": int(t1.rows_fwd),
                                "rows_jac": int(t1.rows_grad), "latent": l1.tolist(), "T_ow": np.asarray(T1).tolist()}
    return out


def fp64_truth(io: dict, n_shape_iters: int, probes: dict | None = None) -> dict:
    """The same steps in fp64 (numpy oracle, oracle/hm_oracle.py): the yardstick for "is the B200 result as close to the exact
    arithmetic as the reference's own fp32 run is" (a 1e-4 comparison between two fp32 runs is only meaningful for
    short runs; the 200-iteration loop and the 39x39 solve amplify rounding).  `probes`: latents whose objective -- the function the
    latent-only LM loop minimises, optimizer.py:345-401: w_recon * sum huber(sdf) + w_codereg * |z|^2 -- is evaluated in fp64."""
    from oracle import hm_oracle as O
    mm, _ = _torch_mm()
    O.set_matmul(mm)
    W, b, _ = load_weights()
    dec = O.DecoderOracle(W, b, (4,), np.float64)
    cfg = copy.deepcopy(WILD_CFG)
    out = {}
    if n_shape_iters:
        cfg["opt"]["converge"]["max_iter"] = n_shape_iters
        lat = io["init_lat"].astype(np.float64).copy()
        tr = O.OptTrace()
        O.shape_opt_deepsdf(dec, cfg, lat, io["T_ow"].astype(np.float64), io["points_w"], trace=tr)
        out["shape_latent"] = lat.tolist()
        for k in (5, 30):
            if len(tr.latent) >= k:
                out[f"shape_latent{k}"] = tr.latent[k - 1].tolist()
        out["shape_states"] = [l.tolist() for l in tr.latent[:REPLAY_STEPS]]       # the state after each of the first iterations (step replay)
        T = io["T_ow"].astype(np.float64)
        pts_o = io["points_w"].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        t_h, w_r, w_c = float(cfg["opt"]["recon"]["robust_th_m"]), float(cfg["opt"]["weight"]["w_recon"]), float(cfg["opt"]["weight"]["w_codereg"])

        def objective(z):
            z = np.asarray(z, np.float64).reshape(-1)
            r = np.abs(np.asarray(O.compute_sdf_loss(dec, z, pts_o, cfg["opt"]["scale_on"])[0], np.float64).reshape(-1))
            return float(w_r * np.where(r <= t_h, r * r, 2 * t_h * r - t_h * t_h).sum() + w_c * (z * z).sum())
        out["objective"] = {"fp64_200": objective(lat), **{k: objective(v) for k, v in (probes or {}).items()}}
    if "rd" in io:
        cfg["opt"]["converge"]["max_iter"] = 1
        tr = O.OptTrace()
        l1 = io["j_init_lat"].astype(np.float64).copy()
        O.shape_pose_joint_opt(dec, cfg, l1, io["j_init_T"].astype(np.float64), io["rd"], io["j_points_w"], 0.08, False, trace=tr)
        out["joint_it0"] = {"H": tr.H[0].tolist(), "b": tr.b[0].tolist(), "rows_fwd": int(tr.rows_fwd), "rows_jac": int(tr.rows_grad)}
    return out


def save_io(path, io):
    flat = {k: v for k, v in io.items() if k != "rd"}
    if "rd" in io:
        flat["rd_n"] = np.int32(len(io["rd"]["T_wc"]))
        for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg"):
            for i, a in enumerate(io["rd"][k]):
                flat[f"rd_{k}_{i}"] = np.asarray(a, np.float32)
    np.savez(path, **flat)


def load_io(path):
    with np.load(path) as z:
        io = {k: z[k] for k in z.files if not k.startswith("rd_")}
        if "rd_n" in z.files:
            n = int(z["rd_n"])
            io["rd"] = {k: [z[f"rd_{k}_{i}"] for i in range(n)] for k in ("T_wc", "rays_fg", "rays_bg", "depth_fg", "depth_bg")}
    return io


def synth_io_cpu(seed=0):
    """Inputs for the stand-alone reference arm without touching the GPU: any 2048 points in the object cube cost the
    same FLOPs (the decoder-only loop's work does not depend on the data)."""
    g = np.random.default_rng(seed)
    _, _, codes = load_weights()
    return {"points_w": ((g.random((N_PTS, 3)) * 2 - 1) * 0.045).astype(np.float32), "T_ow": np.eye(4, dtype=np.float32),
            "init_lat": codes.mean(0).astype(np.float32)}


def run_reference(args):
    """`--impl reference` / `--impl reference-cuda`: the reference's own implementation alone, same metric and config.
    A step is a bounded sample of the workload: ONE fruit x HM_BENCH_REF_ITERS (default 100) of the 200 iterations."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    device = "cuda" if args.impl == "reference-cuda" else "cpu"
    n_it = int(os.environ.get("HM_BENCH_REF_ITERS", "100"))
    io = load_io(args.io) if args.io else synth_io_cpu()
    if args.io:                                    # child of the b200 arm: one full unit + the joint sample, states returned
        res = cpu_arm(io, N_ITERS, int(os.environ.get("HM_BENCH_REF_JOINT_ITERS", "10")), True, device)
        print("HM_CPU_ARM " + json.dumps(res))
        return
    try:
        for _ in range(min(args.warmup, 1)):
            cpu_arm(io, 4, 0, False, device)
        times, info = [], None
        for _ in range(args.steps):
            info = cpu_arm(io, n_it, 0, False, device)
            times.append(info["shape_s"])
    except Exception as e:                          # e.g. reference-cuda without a GPU or without baseline/_ref
        print(json.dumps({"impl": args.impl, "unavailable": f"{type(e).__name__}: {e}"[:300]}))
        return
    step_s = statistics.median(times)
    value = 1.0 / (step_s * N_ITERS / n_it)
    sample = f"{info['what']}: shape_opt_deepsdf, 1 fruit x {N_PTS} pts x {n_it} of {N_ITERS} iterations per step (median of {len(times)} steps), scaled x{N_ITERS / n_it:g}"
    out = {"impl": args.impl, "metric": METRIC, "value": value, "unit": "fruits/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "64 fruits x 2048 pts x 200 LM iters, decoder-only (shape_opt_deepsdf)",
                      "sample": f"1 fruit x {N_PTS} pts x {n_it} of 200 iterations per step, scaled x{N_ITERS / n_it:g}",
                      "step_spread": {"min_s": min(times), "max_s": max(times)}},
           "cpu_baseline": {"value": value, "unit": "fruits/s", "cores": info["cores"], "host_cores": os.cpu_count(), "kind": info["kind"],
                            "sample": sample},
           "e2e": {"value": value, "unit": "fruits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def run_cpu_child(io: dict) -> dict | None:
    """The CPU baseline leg of the b200 arm: a child process (own BLAS thread settings, `.cuda()` neutralised there only)."""
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "io.npz")
        save_io(path, io)
        env = dict(os.environ)
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            env[v] = str(os.cpu_count())
        env["CUDA_VISIBLE_DEVICES"] = ""
        for v in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
            env.pop(v, None)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--io", path], env=env,
                           capture_output=True, text=True, timeout=900)
        for line in r.stdout.splitlines():
            if line.startswith("HM_CPU_ARM "):
                return json.loads(line[len("HM_CPU_ARM "):])
        sys.stderr.write("cpu baseline child failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return None


def peak_tflops():
    peak, src = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
    try:
        pj = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for key in ("bf16_tflops_sustained", "bf16_tflops"):       # the kernel is timed inside a long, power-capped step
            if key in pj and float(pj[key]) > 0:
                return float(pj[key]), f"MEASURED_PEAKS.json {key}"
    except Exception:
        pass
    return peak, src


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--iters", type=int, default=N_ITERS, help=argparse.SUPPRESS)
    ap.add_argument("--io", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--no-joint", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl != "b200":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from hortimapping_b200 import _lib
    from hortimapping_b200.decoder import Decoder, calibration_rows
    from hortimapping_b200.optimizer import Optimizer, PackedBatch, opt_params_from_cfg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    W, b, codes = load_weights()
    dec = Decoder(W, b, device=local)
    g = np.random.default_rng(0)
    dec.calibrate(calibration_rows(codes, 0.15))           # what config_decoder does for a checkpoint directory (ClampingDistance 0.1)
    cfg = copy.deepcopy(WILD_CFG)
    cfg["opt"]["converge"]["max_iter"] = args.iters
    opt = Optimizer(cfg, dec, None, None)
    pts, T_ow, init_lat = make_inputs(dec, codes, seed=7, rank=rank)
    peak, peak_src = peak_tflops()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    params = opt_params_from_cfg(cfg["opt"])

    def timed(step_fn, steps, warmup):
        """W warm-up steps, then K steps between barrier + synchronize, CUDA events, MAX over ranks; decoder launches are
        event-timed inside the region (hm_profile_enable) for the roofline."""
        for _ in range(warmup):
            step_fn()
        sync()
        c0 = dec.counters()
        dec.profile(True)
        sampler = ClockSampler(local)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        ev0.record()
        res = None
        for _ in range(steps):
            res = step_fn()
        ev1.record()
        sync()
        clocks = sampler.stop()
        dec.profile(False)
        c1 = dec.counters()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), {k: c1[k] - c0[k] for k in c1}, clocks, res

    def roofline_of(dc, total_ms, kernel):
        # (a gradient-only row costs the backward half: the same MAC count as a forward row)
        flop = (dc["rows_forward"] + dc["rows_backward"]) * FLOP_PER_FWD_ROW + dc["rows_jacobian"] * FLOP_PER_JAC_ROW
        n_launch, dec_ms = dc["decoder_launches"], dc["decoder_ms"]
        achieved = flop / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else None
        tiles = dc["tiles_forward"] + dc["tiles_jacobian"] + dc["tiles_backward"]
        redone = dc["tiles_redone_forward"] + dc["tiles_redone_jacobian"] + dc["tiles_redone_backward"]
        r = {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
             "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
             "rows_forward": dc["rows_forward"], "rows_jacobian": dc["rows_jacobian"], "rows_gradient_only": dc["rows_backward"],
             "rows_counted": "exact (device-side counters)",
             "algorithmic_flop_per_launch": flop / max(n_launch, 1), "launches": n_launch, "avg_launch_ms": dec_ms / max(n_launch, 1),
             "forward_launches": dc["forward_launches"], "forward_ms": dc["forward_ms"],
             "jacobian_launches": dc["jacobian_launches"], "jacobian_ms": dc["jacobian_ms"],
             "gradient_only_launches": dc["backward_launches"], "gradient_only_ms": dc["backward_ms"],
             "tiles": tiles, "tiles_re_evaluated_with_full_plan": redone,
             "kernel_share_of_step": dec_ms / total_ms if total_ms > 0 else None}
        # `achieved` counts the reference's dense FLOPs (what the reference computes per row).  The engine issues something else:
        # three fp16 products per fp32 product (x3) and only the chunks of the activations that can be non-zero (sparse plan);
        # `issued` is that tensor-core work, from the plan (hm_plan_info) and the exact row / re-evaluated-tile counters
        pi = dec.plan_info()
        ipr = pi["issued_flop_per_row"]
        issued = (dc["rows_forward"] * ipr["sparse_forward"] + dc["rows_jacobian"] * ipr["sparse_jacobian"]
                  + dc["rows_backward"] * (ipr["sparse_jacobian"] - ipr["sparse_forward"])
                  + 64 * (dc["tiles_redone_forward"] * ipr["full_forward"] + dc["tiles_redone_jacobian"] * ipr["full_jacobian"]
                          + dc["tiles_redone_backward"] * (ipr["full_jacobian"] - ipr["full_forward"])))
        r["issued"] = {"what": "fp16 tensor-core FLOP actually issued (3 products per fp32 product, sparse plan, padding, re-evaluated tiles)",
                       "tflops": issued / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else None,
                       "frac_of_peak": issued / (dec_ms * 1e-3) / 1e12 / peak if dec_ms > 0 else None,
                       "flop_per_row": ipr, "alive_chunks_per_layer_of_8": pi["alive_chunks_per_layer"],
                       "dense_plan_would_issue_x": (dc["rows_forward"] * ipr["full_forward"] + dc["rows_jacobian"] * ipr["full_jacobian"]
                                                    + dc["rows_backward"] * (ipr["full_jacobian"] - ipr["full_forward"])) / issued if issued else None}
        tfile = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tfile):       # one ncu --set full capture of this kernel (static file, named so it can go stale visibly)
            tj = json.load(open(tfile))
            r["traffic"] = tj.get("dram_bytes_per_launch")
            r["ncu_capture"] = {k: tj.get(k) for k in ("tensor_pipe_active_pct", "source", "kernel", "rows_per_launch")}
        return r

    # ================================================================== headline: decoder-only loop, device-resident
    lat0 = torch.from_numpy(init_lat).to(dev)
    T0 = torch.from_numpy(T_ow).to(dev)
    gathered = torch.empty(world * N_FRUITS, 49, device=dev)
    pk = PackedBatch([p for p in pts], None, 0, np.zeros(N_FRUITS, np.float32), np.zeros(N_FRUITS, bool))

    def step_device():
        lat, T = lat0.clone(), T0.clone()
        iters, status = opt._run(pk, lat, T, params)
        rec = torch.cat([lat, T.reshape(N_FRUITS, 16), iters.float().reshape(-1, 1)], 1)
        if world > 1:
            dist.all_gather_into_tensor(gathered, rec)
        else:
            gathered.copy_(rec)
        flush.zero_()
        return lat, iters, status

    total_ms, dc, clocks, (lat_out, iters_out, status_out) = timed(step_device, args.steps, args.warmup)
    assert int(iters_out.min().item()) == args.iters, "a fruit exited early: the workload is not fixed"
    value = world * N_FRUITS * args.steps / (total_ms / 1e3) * (args.iters / N_ITERS)   # --iters != 200 is a debug aid only
    roofline = roofline_of(dc, total_ms, "tc_decoder_kernel<true> (fused DeepSDF forward + input gradient)")
    launches = dc["kernel_launches"]

    # ---- end-to-end arm: C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region
    h_lat = torch.from_numpy(init_lat).pin_memory()
    h_T = torch.from_numpy(T_ow).pin_memory()
    h_pts = torch.from_numpy(pts.reshape(-1, 3)).pin_memory()
    h_it = torch.zeros(N_FRUITS, dtype=torch.int32).pin_memory()
    h_st = torch.zeros(N_FRUITS, dtype=torch.int32).pin_memory()
    offsets = (np.arange(N_FRUITS + 1) * N_PTS).astype(np.int64)
    hb = _lib.FruitBatch()
    hb.n_fruits = N_FRUITS
    hb.h_point_offsets = offsets.ctypes.data
    hb.d_points_w = h_pts.data_ptr()
    hb.d_iter_count, hb.d_status = h_it.data_ptr(), h_st.data_ptr()
    h2d = h_lat.numel() * 4 + h_T.numel() * 4 + h_pts.numel() * 4
    d2h = h_lat.numel() * 4 + h_T.numel() * 4 + h_it.numel() * 4 + h_st.numel() * 4

    def step_host():
        w_lat, w_T = h_lat.clone().pin_memory(), h_T.clone().pin_memory()
        hb.d_latents, hb.d_T_ow = w_lat.data_ptr(), w_T.data_ptr()
        _lib.check(dec._L.hm_optimize_shape_host(dec.handle, C.byref(params), C.byref(hb)), "hm_optimize_shape_host")
        return w_lat

    step_host()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        w_lat = step_host()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * N_FRUITS * args.steps / float(e2e_s.item()) * (args.iters / N_ITERS)
    assert torch.allclose(w_lat.to(dev), lat_out, rtol=0, atol=0), "host and device arms disagree"

    out = {"metric": METRIC, "value": value, "unit": "fruits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{N_FRUITS} fruits/GPU x {N_PTS} pts x {args.iters} LM iters, decoder-only (shape_opt_deepsdf), "
                                  "sweetpepper_32 weights, epsilons 0",
                      "parallelism": f"fruits sharded over {world} GPU(s), one all-gather of 49-float records per step",
                      "l2": "256 MiB buffer written between steps (L2 flush); weights are L2-resident by design within a step",
                      "arithmetic": "fp32 semantics: operands split into fp16 hi+lo, three products (A_hi x W_lo, A_lo x W_hi, A_hi x W_hi) as M=128 "
                                    "cta_group::2 tcgen05 MMAs into one fp32 TMEM accumulator per 2 k-chunks, partials summed in fp32 RN registers; "
                                    "sparse plan: hidden units ordered alive-first by calibration, MMAs on all-zero 64-wide activation chunks dropped, checked per tile (violators re-evaluated with the full plan)"},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": "fruits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "roofline": roofline, "f16_saturated_fruits": int((status_out & 0x40).ne(0).sum().item())}

    # ================================================================== joint record: the full shape + pose loop
    joint_io = {}
    if not args.no_joint:
        fruits = make_joint_inputs(dec, codes, seed=7, rank=rank)
        rds = [fruits[i % N_JOINT_DISTINCT].render_data for i in range(N_JOINT)]
        jpts = [fruits[i % N_JOINT_DISTINCT].points_w for i in range(N_JOINT)]
        jlat0 = torch.from_numpy(np.tile(codes.mean(0).astype(np.float32), (N_JOINT, 1))).to(dev)
        jT0 = torch.eye(4, device=dev).repeat(N_JOINT, 1, 1).contiguous()
        jpk = PackedBatch(jpts, rds, cfg["opt"]["render"]["n_frame"], np.full(N_JOINT, 0.08, np.float32), np.zeros(N_JOINT, bool))
        jgathered = torch.empty(world * N_JOINT, 49, device=dev)

        def jstep_device():
            lat, T = jlat0.clone(), jT0.clone()
            iters, status = opt._run(jpk, lat, T, params)
            rec = torch.cat([lat, T.reshape(N_JOINT, 16), iters.float().reshape(-1, 1)], 1)
            if world > 1:
                dist.all_gather_into_tensor(jgathered, rec)
            else:
                jgathered.copy_(rec)
            flush.zero_()
            return lat, T, iters, status

        j_steps = max(1, min(args.steps, 2))
        j_ms, jdc, jclocks, (jlat, jT, jit, jst) = timed(jstep_device, j_steps, 1)
        assert int(jit.min().item()) == args.iters, "joint: a fruit exited early: the workload is not fixed"
        j_value = world * N_JOINT * j_steps / (j_ms / 1e3) * (args.iters / N_ITERS)
        # e2e through hm_optimize_joint_host (host buffers in, results out)
        hp = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        hj = {"lat": hp(np.tile(codes.mean(0).astype(np.float32), (N_JOINT, 1))), "T": hp(np.tile(np.eye(4, dtype=np.float32), (N_JOINT, 1, 1))),
              "pts": hp(jpk.points), "Twc": hp(jpk.T_wc), "rays": hp(jpk.rays), "dobs": hp(jpk.depth_obs),
              "it": torch.zeros(N_JOINT, dtype=torch.int32).pin_memory(), "st": torch.zeros(N_JOINT, dtype=torch.int32).pin_memory()}
        jb = _lib.FruitBatch()
        jb.n_fruits = N_JOINT
        jb.d_points_w, jb.h_point_offsets = hj["pts"].data_ptr(), jpk.point_offsets.ctypes.data
        jb.h_frame_offsets, jb.d_T_wc = jpk.frame_offsets.ctypes.data, hj["Twc"].data_ptr()
        jb.h_ray_offsets, jb.h_n_fg = jpk.ray_offsets.ctypes.data, jpk.n_fg.ctypes.data
        jb.d_rays, jb.d_depth_obs = hj["rays"].data_ptr(), hj["dobs"].data_ptr()
        jb.h_cube_radius, jb.h_pose_known = jpk.cube_radius.ctypes.data, jpk.pose_known.ctypes.data
        jb.d_iter_count, jb.d_status = hj["it"].data_ptr(), hj["st"].data_ptr()
        j_h2d = sum(hj[k].numel() * 4 for k in ("lat", "T", "pts", "Twc", "rays", "dobs"))
        j_d2h = sum(hj[k].numel() * 4 for k in ("lat", "T", "it", "st"))

        def jstep_host():
            w_lat, w_T = hj["lat"].clone().pin_memory(), hj["T"].clone().pin_memory()
            jb.d_latents, jb.d_T_ow = w_lat.data_ptr(), w_T.data_ptr()
            _lib.check(dec._L.hm_optimize_joint_host(dec.handle, C.byref(params), C.byref(jb)), "hm_optimize_joint_host")
            return w_lat

        sync()
        t0 = time.perf_counter()
        wj = jstep_host()
        torch.cuda.synchronize()
        je2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(je2e_s, op=dist.ReduceOp.MAX)
        j_e2e = world * N_JOINT / float(je2e_s.item()) * (args.iters / N_ITERS)
        n_it_total = args.iters * j_steps
        out["joint"] = {
            "metric": "fruits/sec (200 iters, 2048 pts + 10 frames x 400 rays x 30 samples), shape_pose_joint_opt",
            "value": j_value, "unit": "fruits/s", "steps": j_steps, "warmup": 1, "ms_per_step": j_ms / j_steps,
            "config": {"workload": f"{N_JOINT} fruits/GPU ({N_JOINT_DISTINCT} distinct synthetic fruits x {N_JOINT // N_JOINT_DISTINCT}) x (10 frames x (200 fg + 200 bg) rays x 30 samples "
                                   f"+ {N_PTS} pts) x {args.iters} LM iters, Sim(3) + latent (optimizer.py:28-302), configs/wild_pepper.yaml, epsilons 0, "
                                   "20 % leaf-occluded background rays"},
            "e2e": {"value": j_e2e, "unit": "fruits/s", "h2d_bytes_per_step": j_h2d, "d2h_bytes_per_step": j_d2h, "steps": 1},
            "host_equals_device": bool(torch.equal(wj.to(dev), jlat)),
            "roofline": roofline_of(jdc, j_ms, "tc_decoder_kernel<0> (in-sphere ray samples: forward, ReLU bits kept) + <1> (observed points: forward + gradient) + "
                                                  "<2> (in-band samples: gradient only, from the kept bits)"),
            "rows_per_fruit_iteration": {"forward": jdc["rows_forward"] / (N_JOINT * n_it_total), "forward_plus_gradient": jdc["rows_jacobian"] / (N_JOINT * n_it_total),
                                         "gradient_only": jdc["rows_backward"] / (N_JOINT * n_it_total)},
            "clocks": jclocks, "gpu_launches": jdc["kernel_launches"], "status_bits_seen": sorted({hex(int(x)) for x in jst.cpu().tolist()})}
        f0 = fruits[0]
        joint_io = {"j_init_lat": f0.init_latent, "j_init_T": f0.init_T_ow, "j_points_w": f0.points_w, "rd": f0.render_data}
        # iteration 0 of fruit 0 on the device for the in-run parity check (single-fruit call): the LM system, the exact row counts
        # and the state after the step
        c_a = dec.counters()
        l1 = torch.from_numpy(f0.init_latent.copy()).to(dev).reshape(1, 32)
        T1 = torch.from_numpy(f0.init_T_ow.copy()).to(dev).reshape(1, 4, 4)
        opt.shape_pose_joint_opt_batch(l1, T1, [f0.render_data], [f0.points_w], 0.08, False, max_iter=1)
        gH, gb, gdx = (t.cpu().numpy()[0] for t in opt.last_system(1))
        c_b = dec.counters()
        g_it0 = {"H": gH, "b": gb, "dx": gdx, "latent": l1[0].cpu().numpy(), "T_ow": T1[0].cpu().numpy(),
                 "rows_fwd": c_b["rows_forward"] - c_a["rows_forward"], "rows_jac": (c_b["rows_jacobian"] - c_a["rows_jacobian"]) + (c_b["rows_backward"] - c_a["rows_backward"])}
        # the same step with the library's fp32 CUDA-core decoder (HM_ENGINE_SIMT, an independent device implementation of the MLP):
        # separates the loss / normal-equation arithmetic from the rounding of the tensor-core decoder
        dec.set_engine("simt")
        try:
            l2 = torch.from_numpy(f0.init_latent.copy()).to(dev).reshape(1, 32)
            T2 = torch.from_numpy(f0.init_T_ow.copy()).to(dev).reshape(1, 4, 4)
            opt.shape_pose_joint_opt_batch(l2, T2, [f0.render_data], [f0.points_w], 0.08, False, max_iter=1)
            sH, sb, _ = (t.cpu().numpy()[0] for t in opt.last_system(1))
        finally:
            dec.set_engine("tc")

    # ================================================================== CPU baseline + in-run parity (rank 0, N = 1)
    parity_fail = None
    if rank == 0 and world == 1 and not args.no_cpu:
        io = {"points_w": pts[0], "T_ow": T_ow[0], "init_lat": init_lat[0]}
        io.update(joint_io)
        gpu_states = {}
        if args.iters == N_ITERS:
            for k in (5, 30):                       # fruit 0 alone, k iterations from the same start
                lk = torch.from_numpy(init_lat[:1].copy()).to(dev)
                opt.shape_opt_deepsdf_batch(lk, torch.from_numpy(T_ow[:1].copy()).to(dev), [pts[0]], max_iter=k)
                gpu_states[k] = lk[0].cpu().numpy()
            io["probe_b200_30"], io["probe_b200_200"] = gpu_states[30], lat_out[0].cpu().numpy()
        res = run_cpu_child(io)
        if res is not None:
            cv = 1.0 / res["shape_s"]
            out["cpu_baseline"] = {"value": cv, "unit": "fruits/s", "cores": res["cores"], "host_cores": os.cpu_count(), "kind": res["kind"],
                                   "sample": f"{res['what']}: shape_opt_deepsdf on fruit 0 of this run, {N_PTS} pts x {res['shape_iters']} iterations "
                                             f"({res['shape_s']:.1f} s) = one whole unit of the workload"}
            if args.iters == N_ITERS and "shape_states" in res.get("fp64", {}):
                # What can be held to north_star's 1e-4, and what cannot (scripts/diag_parity.py, profiles/r02c_diag_parity.txt):
                # the input gradient of a ReLU network is piecewise constant, so a hidden unit whose pre-activation is within rounding
                # of zero switches sides between two correct evaluations and moves that row's gradient by a few per cent -- an entry
                # of H by ~2e-5, b (which is at its cancellation floor after ~4 iterations) by up to 1e-2, the step by up to 1e-4 of the
                # state.  The loop then amplifies the difference (it never settles: |dx| stays ~1e-4), and every pair of correct runs
                # -- the reference in fp32 and fp64, numpy fp32, this library's two engines -- ends up ~1e-3 apart.  Therefore:
                # (1) GATE: each of the first 5 LM steps is replayed on the device from the fp64 run's own state and must land
                #     within 1e-4 of the fp64 step (typical 1.4e-6; 2.5e-5 when a row crosses a kink);
                # (2) GATE: the 30- and 200-iteration runs must reach the reference's objective (fp64 evaluation) within 1e-2
                #     (the reference's own fp32 / fp64 runs differ by ~2e-3 there);
                # (3) reported: trajectory distances after 5 / 30 / 200 iterations next to the reference's own fp32-vs-fp64 distance.
                st64 = [np.asarray(v, np.float32) for v in res["fp64"]["shape_states"]]
                e_steps = []
                for i in range(len(st64)):
                    li = torch.from_numpy((init_lat[0] if i == 0 else st64[i - 1]).copy()).to(dev).reshape(1, 32)
                    opt.shape_opt_deepsdf_batch(li, torch.from_numpy(T_ow[:1].copy()).to(dev), [pts[0]], iter_offset=i, max_iter=1)
                    e_steps.append(rel_err(li[0].cpu().numpy(), res["fp64"]["shape_states"][i]))
                e_replay = max(e_steps)
                e5 = rel_err(gpu_states[5], res["fp64"]["shape_latent5"])
                e5_ref, e5_ref_64 = rel_err(gpu_states[5], res["shape_latent5"]), rel_err(res["shape_latent5"], res["fp64"]["shape_latent5"])
                e30 = rel_err(gpu_states[30], res["shape_latent30"])
                g200 = lat_out[0].cpu().numpy()
                e_g_ref, e_g_64 = rel_err(g200, res["shape_latent"]), rel_err(g200, res["fp64"]["shape_latent"])
                e_ref_64 = rel_err(res["shape_latent"], res["fp64"]["shape_latent"])
                obj = res["fp64"].get("objective", {})
                o_rel = {k: abs(obj[f"b200_{k}"] - obj[f"reference_{k}"]) / obj[f"reference_{k}"] for k in ("30", "200")
                         if f"b200_{k}" in obj and f"reference_{k}" in obj}
                ok_obj = bool(o_rel) and all(v <= OBJECTIVE_TOL for v in o_rel.values())
                out["parity"] = {"what": "latent of fruit 0, B200 vs the same algorithm in fp64 and vs the CPU baseline run above (max-norm relative); objective = "
                                         "w_recon * sum huber(sdf) + w_codereg * |z|^2 of the final latent, evaluated in fp64",
                                 "step_replay": {"what": f"each of the first {len(st64)} LM steps from the fp64 run's own state vs the fp64 step",
                                                 "per_step": e_steps, "max_rel": e_replay, "tol": PARITY_TOL, "ok": bool(e_replay <= PARITY_TOL)},
                                 "after_5_iterations": {"b200_vs_fp64": e5, "b200_vs_reference_fp32": e5_ref, "reference_fp32_vs_fp64": e5_ref_64},
                                 "after_30_iterations": {"b200_vs_reference_fp32": e30, "objective_rel_diff": o_rel.get("30")},
                                 "after_200_iterations": {"b200_vs_reference_fp32": e_g_ref, "b200_vs_fp64": e_g_64, "reference_fp32_vs_fp64": e_ref_64,
                                                          "objective_rel_diff": o_rel.get("200")},
                                 "objective_fp64": obj, "objective_tol": OBJECTIVE_TOL,
                                 "criterion": "every replayed step <= 1e-4 of the fp64 step; objective after 30 / 200 iterations within 1e-2 of the reference's; "
                                              "trajectory distances are reported, not gated (ReLU-kink crossings are amplified by a loop that never settles: "
                                              "any two correct runs are ~1e-3 apart after ~12 iterations)",
                                 "max_rel": e_replay, "tol": PARITY_TOL, "ok": bool(e_replay <= PARITY_TOL and ok_obj)}
                if not out["parity"]["ok"]:
                    parity_fail = f"headline parity: step replay {e_replay:.3e} (tol {PARITY_TOL:g}), objective rel diff {o_rel} (tol {OBJECTIVE_TOL:g})"
            if "joint" in out and "joint_s" in res:
                jv = 1.0 / (res["joint_s"] * N_ITERS / res["joint_iters"])
                out["joint"]["cpu_baseline"] = {"value": jv, "unit": "fruits/s", "cores": res["cores"], "host_cores": os.cpu_count(), "kind": res["kind"],
                                                "sample": f"{res['what']}: shape_pose_joint_opt on fruit 0 of this run, {res['joint_iters']} of {N_ITERS} iterations "
                                                          f"({res['joint_s']:.1f} s), scaled x{N_ITERS / res['joint_iters']:g} (extrapolated)"}
                if "joint_it0" in res:
                    r0 = res["joint_it0"]
                    # Sample membership is decided by hard thresholds (in-sphere, |sdf| < th, de_do > 1e-6, occlusion): when the device
                    # and the CPU run select the same NUMBER of rows no sample flipped and H, b must agree to 1e-4; a flipped
                    # sample changes H, b by up to its own weight (measured <= 5e-3, tests/test_gpu_optimizer.py), which is then allowed.
                    flips = abs(int(r0["rows_fwd"]) - int(g_it0["rows_fwd"])) + abs(int(r0["rows_jac"]) - int(g_it0["rows_jac"]))
                    errs = {k: rel_err(g_it0[k], r0[k]) for k in ("H", "b", "dx", "latent", "T_ow")}
                    t64 = res.get("fp64", {}).get("joint_it0")
                    vs64 = {}
                    if t64:
                        vs64 = {"b200_H": rel_err(g_it0["H"], t64["H"]), "b200_b": rel_err(g_it0["b"], t64["b"]),
                                "b200_fp32_engine_H": rel_err(sH, t64["H"]), "b200_fp32_engine_b": rel_err(sb, t64["b"]),
                                "reference_fp32_H": rel_err(r0["H"], t64["H"]), "reference_fp32_b": rel_err(r0["b"], t64["b"])}
                    # Gates.  (a) everything but the tensor-core decoder's rounding -- ray sampling, compositing, Jacobian chain rule,
                    # normal equations -- is held to 1e-4 through the fp32-engine run.  (b) The tensor-core run is held to 5e-4: the
                    # input gradient of a ReLU network is piecewise constant, a hidden unit within rounding of zero switches sides
                    # between two correct evaluations and moves that row's gradient by O(10 %) (tests/test_gpu_decoder.py
                    # assert_jac_close); one such row among the ~2000 high-weight in-band samples moves an entry of H by ~1e-4.
                    # (c) A sample that flips membership (row counts differ) changes H, b by up to its own weight.
                    e_simt = max(rel_err(sH, r0["H"]), rel_err(sb, r0["b"]))
                    if flips == 0:
                        ok = e_simt <= PARITY_TOL and errs["H"] <= 5e-4 and errs["b"] <= 5e-4
                    else:
                        ok = errs["H"] <= 5e-3 and errs["b"] <= 2e-2
                    errs["H_fp32_engine"], errs["b_fp32_engine"] = rel_err(sH, r0["H"]), rel_err(sb, r0["b"])
                    out["joint"]["parity"] = {"what": "iteration 0 of fruit 0 from the same start, B200 vs the CPU baseline: normal equations H, b (max-norm relative), "
                                                      "solve dx, state after the step; rows = decoder rows selected by the hard thresholds (equal counts = no sample "
                                                      "flipped membership). Gate with equal counts: <= 1e-4 with the library's fp32 CUDA-core decoder engine, <= 5e-4 with "
                                                      "the tensor-core engine (a ReLU unit within rounding of zero flips one row's gradient); vs_fp64 holds all three "
                                                      "against the same step in fp64. dx / state go through a 39x39 solve with cond ~1e5 (reference: fp32 "
                                                      "torch.inverse, here: fp64 elimination) and are reported; later iterations are compared in tests/ by step replay from "
                                                      "the reference's own states (SURVEY.md 7.4)",
                                              "rows_forward": {"b200": int(g_it0["rows_fwd"]), "cpu": int(r0["rows_fwd"])},
                                              "rows_forward_plus_gradient": {"b200": int(g_it0["rows_jac"]), "cpu": int(r0["rows_jac"])},
                                              "membership_flips": flips, **{"rel_" + k: v for k, v in errs.items()}, "vs_fp64": vs64,
                                              "max_rel": max(errs["H"], errs["b"]), "tol": 5e-4 if flips == 0 else 5e-3, "ok": bool(ok)}
                    if not ok:
                        parity_fail = f"joint parity H {errs['H']:.3e} b {errs['b']:.3e} (vs fp64: {vs64})"
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if parity_fail:
        sys.stderr.write("bench.py: PARITY FAILURE: " + parity_fail + "\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
