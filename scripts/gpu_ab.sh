#!/bin/bash
# Quick GPU iteration loop for decoder-kernel changes: parity first (stop on failure), then speed.
#   HM_EXTRA_NVCC_FLAGS=-DHM_TC_COUNTERS python -c "from hortimapping_b200 import build; build.build_library(force=True)"
# beforehand makes scratch/dbg_wait.py print the per-role wait-cycle breakdown.
mkdir -p gpurun_out
( timeout -k 5 150 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/ab_pytest_decoder.log
tail -3 gpurun_out/ab_pytest_decoder.log
if ! grep -q passed gpurun_out/ab_pytest_decoder.log || grep -q failed gpurun_out/ab_pytest_decoder.log; then echo "decoder tests did not pass: stopping"; exit 1; fi
( timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
( timeout -k 5 200 python scratch/dbg_wait.py 2>&1 | tail -5 ) > gpurun_out/dbg_wait.log; cat gpurun_out/dbg_wait.log
( timeout -k 5 400 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 ) > gpurun_out/bench.log; cat gpurun_out/bench.log
