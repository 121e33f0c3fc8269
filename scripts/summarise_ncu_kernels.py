#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` capture of scripts/ncu_target.py (every kernel of the library):
duration, DRAM traffic, achieved HBM GB/s and tensor-pipe activity against the measured peaks (MEASURED_PEAKS.json).

    python scripts/summarise_ncu_kernels.py gpurun_out/prof_all.ncu-rep profiles/r02a_ncu_kernels.txt
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3,
         "nsecond": 1e-9, "second": 1.0}


def val(r, name, default=0.0):
    if name not in col:
        return default
    try:
        return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1.0)
    except ValueError:
        return default


try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = float(peaks["hbm_gbs"])
except Exception:
    hbm = 6650.0
agg = {}
for r in data:
    name = r[col["Kernel Name"]]
    a = agg.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "tp": 0.0, "sm": 0.0, "regs": 0, "grid": "", "block": "", "smem": 0, "l2": 0.0})
    a["n"] += 1
    a["t"] += val(r, "gpu__time_duration.sum")
    a["rd"] += val(r, "dram__bytes_read.sum")
    a["wr"] += val(r, "dram__bytes_write.sum")
    a["tp"] += val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    a["sm"] += val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    a["l2"] += val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    a["regs"] = int(val(r, "launch__registers_per_thread"))
    a["grid"], a["block"] = r[col["launch__grid_size"]] if "launch__grid_size" in col else "", r[col["launch__block_size"]] if "launch__block_size" in col else ""
    a["smem"] = int(val(r, "launch__shared_mem_per_block_dynamic") + val(r, "launch__shared_mem_per_block_static"))
with open(out, "w") as o:
    o.write("# ncu --set full --clock-control none --profile-from-start off python scripts/ncu_target.py  (one launch window, values averaged per launch;\n")
    o.write(f"# under ncu every launch runs alone and cold: use durations as shares.)  HBM peak = {hbm:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs, measured)\n")
    o.write("%-64s %3s %10s %10s %10s %9s %7s %7s %6s %5s %12s %6s %7s\n" % ("kernel", "n", "time us", "dram rd MB", "dram wr MB", "GB/s", "%HBM", "tensor%", "sm%", "l2%", "grid x block", "regs", "smem B"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        n = a["n"]
        t = a["t"] / n
        gbs = (a["rd"] + a["wr"]) / n / t / 1e9 if t > 0 else 0
        o.write("%-64s %3d %10.1f %10.3f %10.3f %9.1f %7.2f %7.2f %6.1f %5.1f %12s %6d %7d\n" % (
            name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")[:64], n, t * 1e6, a["rd"] / n / 1e6, a["wr"] / n / 1e6, gbs, 100 * gbs / hbm,
            a["tp"] / n, a["sm"] / n, a["l2"] / n, f'{a["grid"]}x{a["block"]}', a["regs"], a["smem"]))
print(open(out).read())
