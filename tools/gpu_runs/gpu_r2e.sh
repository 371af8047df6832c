#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxel2obj_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests_v2o.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests_v2o.log
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/${tag}_v2o_1024.json 2> gpurun_out/${tag}_v2o_1024.err
timeout 600 python tools/bench_voxel2obj.py --size 2048 --steps 3 > gpurun_out/${tag}_v2o_2048.json 2> gpurun_out/${tag}_v2o_2048.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gauss32|approx_|exact_list|select_|nms_|bitonic|finish_rows|sort_prepare' -c 300 --csv --log-file gpurun_out/${tag}_v2o_2048_launches.csv \
    python tools/bench_voxel2obj.py --size 2048 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu1.log 2>&1
exit 0
