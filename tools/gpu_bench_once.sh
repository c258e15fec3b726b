#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['bound'], d['roofline']['limited_by'][:30], d['roofline']['frac'])"
tail -2 gpurun_out/r02_bench.err
