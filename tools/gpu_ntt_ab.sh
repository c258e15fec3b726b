#!/bin/bash
mkdir -p gpurun_out
echo "== kernel parity (persistent/per-polynomial combos)"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "persistent_and_per_polynomial" 2>&1 | tail -3
echo "== ntt_bench 8192: inverse persistent (default), forward per-polynomial"; timeout 120 python tools/ntt_bench.py 8192 5 16384
echo "== ntt_bench 8192: both persistent"; CNHE_NTT_WS_FWD=1 timeout 120 python tools/ntt_bench.py 8192 5 16384
echo "== ntt_bench 8192: none persistent"; CNHE_NTT_WS=0 timeout 120 python tools/ntt_bench.py 8192 5 16384
echo "== ntt_bench 4096"; timeout 120 python tools/ntt_bench.py 4096 3 32768
run() { python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value'], d['roofline']['frac']); print(d['roofline']['families_ms_per_step'])"; }
echo "== bench default"; run
echo "== bench forward persistent too"; CNHE_NTT_WS_FWD=1 run
