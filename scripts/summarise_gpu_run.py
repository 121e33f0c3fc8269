#!/usr/bin/env python
"""Turn the scratch files of one scripts/gpu_round.sh call (gpurun_out/) into the tracked summaries under profiles/.

    python scripts/summarise_gpu_run.py r01a     # -> profiles/r01a_{bench.json,launches.txt,ncu_decoder.txt,...}
"""
import csv, json, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
os.makedirs(PROF, exist_ok=True)


def p(name):
    return os.path.join(PROF, f"{tag}_{name}")


# 1. bench line(s)
for src, dst in (("bench.log", "bench.json"), ("bench_ref.log", "bench_reference.json")):
    f = os.path.join(OUT, src)
    if os.path.exists(f):
        lines = [l for l in open(f) if l.startswith("{")]
        if lines:
            open(p(dst), "w").write(lines[-1])

# 2. launch list: per-kernel count / total device time / share (ncu --metrics gpu__time_duration.sum)
f = os.path.join(OUT, "launches.csv")
if os.path.exists(f):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        agg[r[i_name]][0] += 1
        agg[r[i_name]][1] += float(r[i_val].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(p("launches.txt"), "w") as o:
        o.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n")
        o.write("# command: HM_BENCH_RANDOM_POINTS=1 python bench.py --steps 1 --warmup 1 --iters 40 (first 1200 launches)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("%-100s n=%5d  %10.3f ms  %5.1f%%\n" % (k[:100], v[0], v[1], 100 * v[1] / tot))

# 3. full capture of the decoder kernel: key metrics + stall / opcode summary of the source page
rep = os.path.join(OUT, "prof_decoder.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
            "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster", "smsp__inst_executed.sum",
            "sm__throughput.avg.pct", "lts__t_bytes.sum", "sm__warps_active.avg.pct", "dram__throughput.avg.pct", "launch__shared_mem_per_block_dynamic"]
    with open(p("ncu_decoder.txt"), "w") as o:
        o.write("# ncu --set full --clock-control none --import-source on -k regex:tc_decoder_kernel (values per captured launch)\n")
        for i, h in enumerate(hdr):
            if any(h.startswith(w) or ("." + w) in h for w in want):
                o.write("%-95s %-14s %s\n" % (h, units[i], " | ".join(r[i] for r in rows[2:])))
    i_r = [i for i, h in enumerate(hdr) if h == "dram__bytes_read.sum"][0]
    i_w = [i for i, h in enumerate(hdr) if h == "dram__bytes_write.sum"][0]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tr = [float(r[i_r]) * scale[units[i_r]] + float(r[i_w]) * scale[units[i_w]] for r in rows[2:]]
    i_t = [i for i, h in enumerate(hdr) if h == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
    tp = [float(r[i_t[0]]) for r in rows[2:]] if i_t else []
    json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "tensor_pipe_active_pct": (sum(tp) / len(tp)) if tp else None,
               "source": f"profiles/{tag}_ncu_decoder.txt", "kernel": rows[2][hdr.index("Kernel Name")]},
              open(os.path.join(PROF, "dominant_kernel_traffic.json"), "w"))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join(OUT, "source.csv")
    open(tmp, "w").write(src)
    s = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_summary.py"), tmp, "30"], capture_output=True, text=True).stdout
    open(p("ncu_decoder_source_summary.txt"), "w").write(s)

# 3b. the dominant kernel of the headline workload from the all-kernel capture (scripts/ncu_target.py): tc_decoder_kernel<1, 0> over
#     64 fruits x 2048 points = 131 072 rows per launch, exactly the bench's launch -> DRAM traffic and tensor-pipe activity per launch
f = os.path.join(OUT, "ncu_kernels_raw.csv")
if os.path.exists(f):
    import re
    rows = list(csv.reader(open(f)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}
    val = lambda r, k: float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1.0)
    dom = [r for r in data if re.search(r"tc_decoder_kernel<(\(int\))?1, (\(bool\))?0>", r[col["Kernel Name"]])]
    if dom:
        tmax = max(val(r, "gpu__time_duration.sum") for r in dom)
        big = [r for r in dom if val(r, "gpu__time_duration.sum") > 0.7 * tmax]
        json.dump({"dram_bytes_per_launch": sum(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in big) / len(big),
                   "tensor_pipe_active_pct": sum(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") for r in big) / len(big),
                   "duration_ms_under_ncu": sum(val(r, "gpu__time_duration.sum") for r in big) / len(big) * 1e3,
                   "rows_per_launch": 64 * 2048, "launches_averaged": len(big),
                   "source": f"profiles/{tag}_ncu_kernels.txt (ncu --set full of scripts/ncu_target.py)", "kernel": big[0][col["Kernel Name"]]},
                  open(os.path.join(PROF, "dominant_kernel_traffic.json"), "w"))

# 4. small logs kept verbatim
for name in ("pytest_gpu.log", "dbg_wait.log", "gpu.txt", "extra.log", "ncu_kernels.txt", "ncu_source_fwd.txt", "ncu_source_jac.txt", "ncu_source_bwd.txt",
             "ncu_source_normal_eq.txt", "ncu_source_solve.txt", "sanitizer_memcheck.log", "sanitizer_racecheck.log", "diag_parity.log", "probe_speed.log"):
    f = os.path.join(OUT, name)
    if os.path.exists(f):
        open(p(name.replace(".log", ".txt")), "w").write(open(f).read())
print("wrote", sorted(x for x in os.listdir(PROF) if x.startswith(tag)))
