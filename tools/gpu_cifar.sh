#!/bin/bash
mkdir -p gpurun_out
echo "== bench lola_cifar"; timeout 900 python bench.py --workload lola_cifar --steps 3 --warmup 1 2>gpurun_out/r02_bench_cifar.err | tail -1 > gpurun_out/r02_bench_lola_cifar.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_lola_cifar.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families_ms_per_step'], d['cpu_baseline']['host'])"
tail -3 gpurun_out/r02_bench_cifar.err
