#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mac_umma -c 1 -o gpurun_out/r02_mac_umma_conv python tools/e2e_timeline.py 2 > /dev/null 2>&1
ls -la gpurun_out/r02_mac_umma_conv.ncu-rep
