#!/bin/bash
# Round-end profile pass on one B200: bench line, launch list, ncu --set full of the dominant kernels.
tag=${1:-r01}
mkdir -p gpurun_out
timeout 600 python bench.py 2>gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
C512="python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $C512 > gpurun_out/${tag}_plain512.log 2>&1 || exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches_512.csv $C512 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fused12_kernel -s 1 -c 1 -o gpurun_out/${tag}_prof_fused12 -f $C512 > /dev/null 2>&1
ncu -i gpurun_out/${tag}_prof_fused12.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_fused12_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_strided|gauss_contig|hist1|dense_pass1|dense_pass2" -c 6 -o gpurun_out/${tag}_prof_detect -f $C512 > /dev/null 2>&1
ncu -i gpurun_out/${tag}_prof_detect.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_detect_raw.csv 2>/dev/null
timeout 600 python tools/bench_voxel2obj.py --size 2048 > gpurun_out/${tag}_v2o_2048.json 2>/dev/null
timeout 600 python tools/bench_voxel2obj.py --size 2048 --kind uniform > gpurun_out/${tag}_v2o_2048u.json 2>/dev/null
rm -f gpurun_out/${tag}_prof_fused12.ncu-rep gpurun_out/${tag}_prof_detect.ncu-rep
exit 0
