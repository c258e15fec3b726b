#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "key_switch_mac_full_waves or relin or rotate" 2>&1 | tail -3
echo "== network parity"; timeout 600 python -m pytest tests/test_gpu_network.py -x -q 2>&1 | tail -2
run() { python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value']); print(d['roofline']['families_ms_per_step'])"; }
echo "== bench staged"; run
echo "== bench register kernel"; CNHE_KSMAC_TMA=0 run
