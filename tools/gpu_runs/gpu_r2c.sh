#!/bin/bash
# round-2 GPU pass C: two-tier detection path -- parity tests, cfg4 timing, launch list
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxel2obj_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests_v2o.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests_v2o.log
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/${tag}_v2o_1024.json 2> gpurun_out/${tag}_v2o_1024.err
timeout 300 python tools/bench_voxel2obj.py --size 1024 --kind uniform --steps 3 > gpurun_out/${tag}_v2o_1024u.json 2>> gpurun_out/${tag}_v2o_1024.err
timeout 600 python tools/bench_voxel2obj.py --size 2048 --steps 3 > gpurun_out/${tag}_v2o_2048.json 2> gpurun_out/${tag}_v2o_2048.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_v2o_1024_launches.csv \
    python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 1 > gpurun_out/${tag}_ncu.log 2>&1
exit 0
