#!/bin/bash
mkdir -p gpurun_out
echo "== bench numa=1"; timeout 600 python bench.py --steps 20 --warmup 5 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step','numa')}, d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['roofline']['frac'], d['roofline']['families_ms_per_step'])"
echo "== bench numa=0"; CNHE_NUMA_BIND=0 timeout 600 python bench.py --steps 20 --warmup 5 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets_nonuma.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets_nonuma.json'));print({k:d[k] for k in ('value','ms_per_step','numa')}, d['e2e']['value'])"
tail -3 gpurun_out/r02_bench.err
