#!/bin/bash
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gauss32|approx_|exact_|select_|nms_|bitonic|finish_rows|sort_prepare' -c 300 --csv --log-file gpurun_out/${tag}_v2o_2048_launches.csv \
    python tools/bench_voxel2obj.py --size 2048 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu1.log 2>&1
exit 0
