#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/z2_tests.log 2>&1
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/z2_bench.err | tail -1 > gpurun_out/z2_bench.json
timeout 300 python gpurun_exp6.py > gpurun_out/z2_unet512.log 2>&1
exit 0
