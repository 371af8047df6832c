#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_global_v2o_gpu.py -x -q > gpurun_out/q_tests.log 2>&1
timeout 600 python -m pytest tests/test_voxel2obj_gpu.py -x -q > gpurun_out/q_tests2.log 2>&1
exit 0
