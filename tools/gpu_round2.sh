#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; (timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value'], d['roofline']['frac']); print(d['roofline']['families_ms_per_step'])"
tail -3 gpurun_out/r02_bench.err
echo "== bench KSMAC_CT=1"; CNHE_KSMAC_CT=1 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['families_ms_per_step'])"
echo "== bench KSMAC_CT=2"; CNHE_KSMAC_CT=2 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['families_ms_per_step'])"
