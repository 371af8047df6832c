#!/bin/bash
tag=${1:-un}
mkdir -p gpurun_out
C="python bench.py --model unet_like2 --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
timeout -s KILL 300 $C > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"upcat_blocked_kernel|pool_blocked_kernel|final_blocked_kernel" -s 6 -c 5 -o gpurun_out/${tag}_prof_aux -f $C > gpurun_out/${tag}_ncu.log 2>&1
exit 0
