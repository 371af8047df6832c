#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/n_tests.log 2>&1
timeout 300 python gpurun_exp6.py > gpurun_out/n_unet512.log 2>&1
FPL_NO_KSPLIT=1 timeout 300 python gpurun_exp6.py > gpurun_out/n_unet512_noks.log 2>&1
exit 0
