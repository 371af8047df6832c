#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_network_gpu.py -x -q > gpurun_out/z_tests.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>gpurun_out/z_bench.err | tail -1 > gpurun_out/z_bench.json
FPL_PLAN_LEGACY=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 2 2>gpurun_out/z_bench_legacy.err | tail -1 > gpurun_out/z_bench_legacy.json
timeout 300 python gpurun_exp6.py > gpurun_out/z_unet512.log 2>&1
FPL_PLAN_LEGACY=1 timeout 300 python gpurun_exp6.py > gpurun_out/z_unet512_legacy.log 2>&1
exit 0
