#!/bin/bash
# 2-GPU checks (gpurun --gpus 2): sharded CIFAR correctness, CryptoNets bench with the in-loop score all-gather, CIFAR bench replicas vs row shards
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== shard_check"; timeout 600 $TR --master-port 29511 tools/shard_check.py 9 2>&1 | tail -3
echo "== cryptonets 2 gpu"; timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r02_bench_cryptonets_2gpu.json | cut -c1-900
echo "== cifar 1 gpu"; timeout 900 python bench.py --workload lola_cifar --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_lola_cifar_1gpu.json | cut -c1-1200
echo "== cifar 2 gpu sharded"; timeout 900 $TR --master-port 29513 bench.py --workload lola_cifar --shard-rows --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_lola_cifar_2gpu_shard.json | cut -c1-1200
