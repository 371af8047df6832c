#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 2 > gpurun_out/y_v2o_1024.json 2> gpurun_out/y_v2o_1024.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/y_v2o_1024_launches.csv python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/y_ncu.log 2>&1
timeout 600 python -m pytest tests/test_umma_gpu.py -x -q -k f1 > gpurun_out/y_f1.log 2>&1
exit 0
