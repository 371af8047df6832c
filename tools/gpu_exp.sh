#!/bin/bash
mkdir -p gpurun_out
C512="python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
FPL_GAUSS_CERT=64 timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/l_v2o_1024_cert.json 2> gpurun_out/l_cert.err
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/l_v2o_1024.json 2> gpurun_out/l.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_strided|gauss_contig|dense_pass1|dense_pass2|select_hist" -c 6 -o gpurun_out/l_prof_detect -f $C512 > /dev/null 2>&1
ncu -i gpurun_out/l_prof_detect.ncu-rep --page raw --csv > gpurun_out/l_prof_detect_raw.csv 2>/dev/null
rm -f gpurun_out/l_prof_detect.ncu-rep
exit 0
