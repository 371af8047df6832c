#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i_tests.log 2>&1
timeout 600 python bench.py 2>gpurun_out/i_bench.err | tail -1 > gpurun_out/i_bench.json
timeout 300 python gpurun_exp6.py > gpurun_out/i_unet512.log 2>&1
timeout 900 python tools/bench_voxel2obj.py --size 2048 > gpurun_out/i_v2o_2048.json 2> gpurun_out/i_v2o_2048.err
exit 0
