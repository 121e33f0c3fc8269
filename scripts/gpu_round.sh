#!/bin/bash
# One gpurun call: GPU parity suite, bench line, ncu launch list, one full ncu capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -5 ) > gpurun_out/bench.log
( HM_BENCH_RANDOM_POINTS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --iters 40 2>&1 | tail -3 ) > gpurun_out/ncu_launches.log
( HM_BENCH_RANDOM_POINTS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_decoder_kernel -s 40 -c 2 \
    -f -o gpurun_out/prof_decoder python bench.py --steps 1 --warmup 1 --iters 30 2>&1 | tail -3 ) > gpurun_out/ncu_full.log
( timeout -k 5 600 python scripts/bench_extra.py 2>&1 | grep "^{" ) > gpurun_out/extra.log; cat gpurun_out/extra.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log
