#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; (timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value'], d['roofline']['frac'], d['cpu_baseline']['value']); print(d['roofline']['families_ms_per_step'])"
tail -2 gpurun_out/r02_bench.err
