#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py -x -q > gpurun_out/y2_tests.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>gpurun_out/y2_bench.err | tail -1 > gpurun_out/y2_bench.json
exit 0
