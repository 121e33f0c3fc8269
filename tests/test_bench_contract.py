"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`) prints ONE
JSON line with the agreed keys, on the same metric / unit / config as the B200 arm, and only rank 0 prints under torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, HM_BENCH_REF_ITERS="4", **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "1"],
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_json_line():
    lines = _run({"OMP_NUM_THREADS": "1"})                  # what torchrun exports to every rank
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fruits/sec (200 iters, 2048 pts)" and d["unit"] == "fruits/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    # the unmodified reference when it is staged (/root/reference here, baseline/_ref on the GPU box), else the numpy port
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert ("unmodified reference" in cb["sample"]) if cb["kind"] == "reference" else ("BLAS" in cb["sample"])
    assert d["e2e"] == {"value": d["value"], "unit": "fruits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_falls_back_to_the_port_without_the_reference(monkeypatch):
    """Without /root/reference and baseline/_ref the arm times the oracle port (torch BLAS, all cores) and says so."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import ref_runner
    monkeypatch.setattr(ref_runner, "CANDIDATES", [])
    res = bench.cpu_arm(bench.synth_io_cpu(), 2, 0, False)
    assert res["kind"] == "port" and res["shape_iters"] == 2 and res["shape_s"] > 0 and "BLAS" in res["what"]
    from oracle import hm_oracle as O
    O.set_matmul(None)
