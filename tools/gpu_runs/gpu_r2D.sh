#!/bin/bash
# round-2 ncu evidence for the dominant kernel (same commands as round 1: bench at 512^3)
tag=${1:-r2D}
mkdir -p gpurun_out
timeout 300 python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${tag}_plain512.log 2>&1 || exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_512.csv \
    python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fused12_kernel -s 1 -c 1 -o gpurun_out/${tag}_prof_fused12 -f \
    python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${tag}_ncu_fused.log 2>&1
exit 0
