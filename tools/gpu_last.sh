#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -x -q -k "tensor_cores or randomised or mac_layer or cryptonets or network" 2>&1 | tail -3
