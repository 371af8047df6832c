#!/bin/bash
tag=${1:-r2y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxel2obj_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/${tag}_v2o_1024_imm.json 2> gpurun_out/${tag}_v2o_1024.err
FPL_GAUSS_IMM=0 timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/${tag}_v2o_1024_ur.json 2>> gpurun_out/${tag}_v2o_1024.err
timeout 600 python tools/bench_voxel2obj.py --size 2048 --steps 3 > gpurun_out/${tag}_v2o_2048_imm.json 2> gpurun_out/${tag}_v2o_2048.err
exit 0
