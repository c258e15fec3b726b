#!/bin/bash
mkdir -p gpurun_out
echo "== new randomised tensor-core test"; timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "randomised" 2>&1 | tail -2
echo "== kernel parity with three forward CTAs per SM"; CNHE_NTT_FWD_BLOCKS=3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "cryptonets8192 or lola8192" 2>&1 | tail -3
echo "== ntt_bench 8192 default (2 CTAs/SM)"; timeout 120 python tools/ntt_bench.py 8192 5 16384 | head -1
echo "== ntt_bench 8192 three CTAs/SM"; CNHE_NTT_FWD_BLOCKS=3 timeout 120 python tools/ntt_bench.py 8192 5 16384 | head -1
run() { python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value'], d['roofline']['frac']); print(d['roofline']['families_ms_per_step'])"; }
echo "== bench default"; run
echo "== bench three CTAs/SM"; CNHE_NTT_FWD_BLOCKS=3 run
