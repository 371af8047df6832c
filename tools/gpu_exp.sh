#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1
timeout 300 python gpurun_exp6.py > gpurun_out/o_unet512.log 2>&1
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/o_v2o_1024.json 2> gpurun_out/o.err
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 --kind uniform > gpurun_out/o_v2o_1024u.json 2> gpurun_out/ou.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/o_v2o_1024_launches.csv python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/o_ncu.log 2>&1
exit 0
