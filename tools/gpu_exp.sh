#!/bin/bash
mkdir -p gpurun_out
C512="python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_strided|gauss_contig|hist1|dense_pass1|dense_pass2" -c 6 -o gpurun_out/r01d_prof_detect -f $C512 > /dev/null 2>&1
ncu -i gpurun_out/r01d_prof_detect.ncu-rep --page raw --csv > gpurun_out/r01d_prof_detect_raw.csv 2>/dev/null
rm -f gpurun_out/r01d_prof_detect.ncu-rep
exit 0
