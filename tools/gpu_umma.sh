#!/bin/bash
mkdir -p gpurun_out
echo "== dense + conv parity"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "tensor_cores or mac_layer" 2>&1 | tail -6
echo "== network parity"; timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_wrapper.py -x -q 2>&1 | tail -3
echo "== role profile"; CNHE_UMMA_PROF=1 MS=0 timeout 300 python tools/e2e_timeline.py 2 2>&1 | grep umma | head -6
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['value_two_streams']['value'], d['roofline']['frac']); print(d['roofline']['families_ms_per_step'])"
tail -3 gpurun_out/r02_bench.err
