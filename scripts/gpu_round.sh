#!/bin/bash
# One gpurun call; STEPS selects what runs (default: everything):
#   tests   GPU parity suite            bench   one bench line              launches  ncu launch list of the bench command
#   full    ncu --set full of every kernel (scripts/ncu_target.py), summarised ON THE BOX (the report itself is too big to return)
#   memcheck / racecheck   compute-sanitizer on a subset of the GPU tests     extra     scripts/bench_extra.py
STEPS=${STEPS:-"tests bench launches full memcheck racecheck extra"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then
  ( timeout -k 5 900 python -m pytest tests -m gpu -q -rA 2>&1 | grep -E "passed|failed|PASSED|FAILED|sequence of|iterations replayed" | tail -120 ) > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
fi
if has bench; then
  ( timeout -k 5 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 2>&1 | tail -4 ) > gpurun_out/bench.log; tail -c 700 gpurun_out/bench.log
fi
if has launches; then
  ( HM_BENCH_RANDOM_POINTS=1 timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --iters 20 --no-cpu 2>&1 | tail -3 ) > gpurun_out/ncu_launches.log
fi
if has full; then
  ( timeout -k 5 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/prof_all \
      python scripts/ncu_target.py 2>&1 | tail -3 ) > gpurun_out/ncu_full.log; tail -2 gpurun_out/ncu_full.log
  python scripts/summarise_ncu_kernels.py /tmp/prof_all.ncu-rep gpurun_out/ncu_kernels.txt > /dev/null 2>&1
  ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/ncu_kernels_raw.csv 2>/dev/null
  # the two decoder instantiations share a name prefix: split by launch id (forward-only launches carry <(bool)0>)
  python - <<'PY'
import csv, subprocess, sys, os, re
rows = list(csv.reader(subprocess.run(["ncu", "-i", "/tmp/prof_all.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]; iid, iname = hdr.index("ID"), hdr.index("Kernel Name")
pick = {}
for r in rows[2:]:
    n = r[iname]
    # tc_decoder_kernel<kMode, kRedo> (0 forward, 1 forward + gradient, 2 gradient only): the redo instantiations (second template argument 1) exit at once on these inputs
    key = ("jac" if re.search(r"tc_decoder_kernel<(\(int\))?1, (\(bool\))?0>", n) else "fwd" if re.search(r"tc_decoder_kernel<(\(int\))?0, (\(bool\))?0>", n)
           else "bwd" if re.search(r"tc_decoder_kernel<(\(int\))?2, (\(bool\))?0>", n) else "solve" if "solve_kernel" in n else "normal_eq" if "normal_eq" in n else None)
    if key and key not in pick: pick[key] = r[iid]
for key, lid in pick.items():
    src = subprocess.run(["ncu", "-i", "/tmp/prof_all.ncu-rep", "--page", "source", "--csv", "--launch-skip", lid, "--launch-count", "1"], capture_output=True, text=True).stdout
    open(f"/tmp/src_{key}.csv", "w").write(src)
    s = subprocess.run([sys.executable, "scripts/ncu_source_summary.py", f"/tmp/src_{key}.csv", "25"], capture_output=True, text=True)
    open(f"gpurun_out/ncu_source_{key}.txt", "w").write(s.stdout + s.stderr[-500:])
PY
  ls -la /tmp/prof_all.ncu-rep | tee -a gpurun_out/ncu_full.log
fi
if has memcheck; then
  ( HM_TEST_CAL_ROWS=16384 timeout -k 5 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_optimizer.py -m gpu -q -x \
      -k "sparse_plan or contradict or ragged or replay or batch_equals or mask_reuse or invalid_submap or degenerate" 2>&1 | tail -15 ) > gpurun_out/sanitizer_memcheck.log; tail -4 gpurun_out/sanitizer_memcheck.log
fi
if has racecheck; then
  ( HM_TEST_CAL_ROWS=16384 timeout -k 5 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x \
      -k "ragged and (129 or 1000)" > /tmp/racecheck_full.log 2>&1 )
  ( HM_TEST_CAL_ROWS=16384 timeout -k 5 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_optimizer.py -m gpu -q -x \
      -k "mask_reuse" >> /tmp/racecheck_full.log 2>&1 )
  python scripts/summarise_racecheck.py /tmp/racecheck_full.log > gpurun_out/sanitizer_racecheck.log 2>&1; tail -12 gpurun_out/sanitizer_racecheck.log
fi
if has extra; then
  ( timeout -k 5 600 python scripts/bench_extra.py grid iso nn render_data 2>&1 | grep "^{" ) > gpurun_out/extra.log; cat gpurun_out/extra.log | cut -c1-300
fi
du -sh gpurun_out
