#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_network_gpu.py tests/test_umma_gpu.py -q > gpurun_out/x2_tests.log 2>&1
exit 0
