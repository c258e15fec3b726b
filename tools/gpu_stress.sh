#!/bin/bash
mkdir -p gpurun_out
CNHE_UMMA_PROF=1 timeout 600 python tools/umma_stress.py 40 11 2> gpurun_out/stress.err | tail -45
echo "tcgen05 launches reported:" $(grep -c "\[umma" gpurun_out/stress.err)
grep -v "\[umma" gpurun_out/stress.err | tail -5
