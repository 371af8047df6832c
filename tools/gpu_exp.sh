#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py -x -q -k "pipelined" > gpurun_out/v_tests.log 2>&1
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/v_bench.err | tail -1 > gpurun_out/v_bench.json
exit 0
