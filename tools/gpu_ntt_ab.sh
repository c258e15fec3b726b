#!/bin/bash
mkdir -p gpurun_out
echo "== kernel parity (ws on)"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_wrapper.py -x -q 2>&1 | tail -4
echo "== ntt_bench WS"; timeout 120 python tools/ntt_bench.py 8192 5 16384; timeout 120 python tools/ntt_bench.py 4096 3 32768
echo "== square_bench ws both"; timeout 120 python tools/square_bench.py 845 4
echo "== square_bench ws inv only"; CNHE_NTT_WS_FWD=0 timeout 120 python tools/square_bench.py 845 4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ntt_forward_ws -s 2 -c 1 -o gpurun_out/r02_ws5_fwd python tools/ntt_bench.py 8192 5 16384 > /dev/null 2>&1
