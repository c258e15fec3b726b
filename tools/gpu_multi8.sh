#!/bin/bash
# N-GPU sanity run of the driver's scaling command (gpurun --gpus N): bench under torchrun, then the reference arm under torchrun (rank 0 works, the others exit)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== cryptonets $N gpu"; timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_bench_${N}gpu.err | tail -1 | tee gpurun_out/r02_bench_cryptonets_${N}gpu.json | cut -c1-1000
tail -3 gpurun_out/r02_bench_${N}gpu.err
echo "== reference arm under torchrun"; timeout 600 $TR --master-port 29522 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | grep -v "^\*\|Setting OMP\|^W0\|^$" | tail -2 | cut -c1-600
