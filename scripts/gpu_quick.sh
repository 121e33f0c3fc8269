#!/bin/bash
# One short gpurun call while iterating: decoder parity first (stop on failure), the whole GPU suite, kernel speed, one bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( timeout -k 5 300 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/pytest_decoder.log
tail -5 gpurun_out/pytest_decoder.log
if ! grep -q passed gpurun_out/pytest_decoder.log || grep -q failed gpurun_out/pytest_decoder.log; then echo "decoder tests did not pass: stopping"; exit 1; fi
( timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
( timeout -k 5 200 python scripts/probe_decoder.py speed 2>&1 | tail -4 ) > gpurun_out/probe_speed.log; cat gpurun_out/probe_speed.log
( timeout -k 5 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 2>&1 | tail -4 ) > gpurun_out/bench.log; cat gpurun_out/bench.log
