#!/bin/bash
# A/B of the CTA-pair decoder (HM_TC_PAIR=1, default) against the single-CTA variant: parity first, then speed.
mkdir -p gpurun_out
( timeout -k 5 150 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/ab_pytest_decoder_pair.log
tail -3 gpurun_out/ab_pytest_decoder_pair.log
if ! grep -q passed gpurun_out/ab_pytest_decoder_pair.log || grep -q failed gpurun_out/ab_pytest_decoder_pair.log; then echo "pair decoder tests did not pass: stopping"; exit 1; fi
( timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
( timeout -k 5 200 python scratch/dbg_wait.py 2>&1 | tail -5 ) > gpurun_out/dbg_wait.log; cat gpurun_out/dbg_wait.log
( timeout -k 5 400 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 ) > gpurun_out/bench.log; cat gpurun_out/bench.log
if [ "$1" != "nosingle" ]; then
( HM_TC_PAIR=0 timeout -k 5 150 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/ab_pytest_decoder_single.log
tail -2 gpurun_out/ab_pytest_decoder_single.log
( HM_TC_PAIR=0 timeout -k 5 150 python scratch/dbg_wait.py 2>&1 | tail -5 ) > gpurun_out/dbg_wait_single.log; cat gpurun_out/dbg_wait_single.log
fi
