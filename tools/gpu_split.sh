#!/bin/bash
# N = 16384 on CTA pairs: parity, A/B against the one-CTA kernels, CIFAR timing; plus the host-stall trace of the pipelined e2e loop
mkdir -p gpurun_out
echo "== parity (cifar16384 kernels, both forms)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "cifar16384" 2>&1 | tail -4
echo "== ntt_bench 16384 split"; timeout 120 python tools/ntt_bench.py 16384 8 4096
echo "== ntt_bench 16384 whole"; CNHE_NTT_SPLIT=0 timeout 120 python tools/ntt_bench.py 16384 8 4096
echo "== cifar split"; timeout 300 python tools/cifar_once.py 3
echo "== cifar whole"; CNHE_NTT_SPLIT=0 timeout 300 python tools/cifar_once.py 3
echo "== e2e stall trace"; CNHE_TRACE_SLOW=15 MS=1 timeout 300 python tools/e2e_timeline.py 24 2> gpurun_out/r02_e2e_trace.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step']);[print(r) for r in d['forward_export_import_wait_ms']]"
grep -v "^$" gpurun_out/r02_e2e_trace.err | tail -40
