#!/bin/bash
# N-GPU bench exactly as the driver launches it (one rank per GPU under torch.distributed.run); N = $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt 2>&1
( timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -4 ) > gpurun_out/scale_n$N.log
cat gpurun_out/scale_n$N.log
( timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | tail -3 ) > gpurun_out/scale_ref_n$N.log
cat gpurun_out/scale_ref_n$N.log
( timeout -k 10 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 ) > gpurun_out/scale_n1.log
cat gpurun_out/scale_n1.log
