#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 1 > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gauss32|approx_|exact_list|select_|nms_|bitonic|finish_rows|sort_prepare' -c 400 --csv --log-file gpurun_out/${tag}_v2o_1024_launches.csv \
    python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 1 > gpurun_out/${tag}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gauss32_kernel|approx_pass1_kernel' -c 2 -o gpurun_out/${tag}_prof_v2o -f \
    python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_prof_v2o.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_v2o_raw.csv 2>/dev/null
exit 0
