#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_voxel2obj_gpu.py -x -q > gpurun_out/g_tests.log 2>&1
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/g_v2o_1024.json 2> gpurun_out/g_v2o_1024.err
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 --kind uniform > gpurun_out/g_v2o_1024u.json 2> gpurun_out/g_v2o_1024u.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/g_v2o_1024_launches.csv python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/g_ncu.log 2>&1
exit 0
