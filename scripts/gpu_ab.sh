#!/bin/bash
# Quick GPU iteration loop for decoder-kernel changes: parity first (stop on failure), then speed.
mkdir -p gpurun_out
( timeout -k 5 200 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/ab_pytest_decoder.log
tail -3 gpurun_out/ab_pytest_decoder.log
if ! grep -q passed gpurun_out/ab_pytest_decoder.log || grep -q failed gpurun_out/ab_pytest_decoder.log; then echo "decoder tests did not pass: stopping"; exit 1; fi
( timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
( timeout -k 5 200 python scripts/probe_decoder.py speed 2>&1 | tail -4 ) > gpurun_out/probe_speed.log; cat gpurun_out/probe_speed.log
( timeout -k 5 200 python scripts/probe_decoder.py wait 2>&1 | tail -8 ) > gpurun_out/probe_wait.log; cat gpurun_out/probe_wait.log
( timeout -k 5 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 ) > gpurun_out/bench.log; cat gpurun_out/bench.log
