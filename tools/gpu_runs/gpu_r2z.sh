#!/bin/bash
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxel2obj_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/${tag}_v2o_1024_z64.json 2> gpurun_out/${tag}_v2o_1024.err
FPL_GAUSS_Z64=0 timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/${tag}_v2o_1024_f32.json 2>> gpurun_out/${tag}_v2o_1024.err
timeout 600 python tools/bench_voxel2obj.py --size 2048 --steps 3 > gpurun_out/${tag}_v2o_2048_z64.json 2> gpurun_out/${tag}_v2o_2048.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gauss64z_kernel' -c 1 -o gpurun_out/${tag}_prof_g64 -f \
    python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu2.log 2>&1
exit 0
