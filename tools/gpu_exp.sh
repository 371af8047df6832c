#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_voxel2obj_gpu.py -x -q > gpurun_out/p_tests.log 2>&1
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/p_v2o_1024.json 2> gpurun_out/p.err
exit 0
