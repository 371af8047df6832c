#!/bin/bash
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gauss32_kernel|approx_pass1_kernel|nms_suppress_kernel|approx_ballcheck_kernel|approx_filter_kernel' -c 5 -o gpurun_out/${tag}_prof_v2o -f \
    python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu2.log 2>&1
exit 0
