#!/bin/bash
# one-GPU round job: full GPU test suite, headline bench, LoLa workloads, reference arm, noise trace
mkdir -p gpurun_out
echo "== pytest -m gpu"; (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
echo "== bench cryptonets"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; cut -c1-1500 gpurun_out/r02_bench_cryptonets.json
echo "== bench lola_small"; timeout 600 python bench.py --workload lola_small --steps 20 --warmup 3 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_lola_small.json; cut -c1-1200 gpurun_out/r02_bench_lola_small.json
echo "== bench lola_cifar"; timeout 900 python bench.py --workload lola_cifar --steps 3 --warmup 1 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_lola_cifar.json; cut -c1-1200 gpurun_out/r02_bench_lola_cifar.json
cp profiles/r02_opcounts_*.json gpurun_out/ 2>/dev/null
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_reference.json; cut -c1-1800 gpurun_out/r02_bench_reference.json
tail -5 gpurun_out/r02_bench.err
echo "== noise trace"; timeout 900 python tools/noise_trace.py --extra-prime --out gpurun_out/noise_trace_r02c.json > gpurun_out/r02_noise.log 2>&1; tail -2 gpurun_out/r02_noise.log
