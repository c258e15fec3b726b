#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
echo "== bench cryptonets"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value'], d['roofline']['frac'], d['cpu_baseline']['value']); print(d['roofline']['families_ms_per_step'])"
echo "== bench lola_cifar"; timeout 600 python bench.py --workload lola_cifar --steps 3 --warmup 1 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_lola_cifar.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_lola_cifar.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families_ms_per_step'])"
echo "== bench lola_small"; timeout 600 python bench.py --workload lola_small --steps 20 --warmup 3 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_lola_small.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_lola_small.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families_ms_per_step'])"
tail -2 gpurun_out/r02_bench.err
