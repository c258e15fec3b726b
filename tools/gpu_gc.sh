#!/bin/bash
mkdir -p gpurun_out
show='import json,sys;d=json.loads(sys.stdin.read());print(d["ms_per_step"]);print([r for r in d["forward_export_import_wait_ms"] if r[0]>20 or r[2]>20])'
for mode in default off freeze; do
  echo "== e2e timeline GC=$mode"; GC=$mode MS=1 timeout 300 python tools/e2e_timeline.py 40 2> gpurun_out/gc_$mode.err | python -c "$show"; tail -2 gpurun_out/gc_$mode.err | cut -c1-400
done
echo "== ncu split kernels"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_forward_split -s 4 -c 1 -o gpurun_out/r02_split_fwd python tools/ntt_bench.py 16384 8 4096 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_inverse_split -s 4 -c 1 -o gpurun_out/r02_split_inv python tools/ntt_bench.py 16384 8 4096 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
