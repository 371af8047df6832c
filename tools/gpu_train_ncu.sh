#!/bin/bash
tag=${1:-r3e}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/${tag}_train.log 2>&1
C="python tools/bench_train.py --precision tf32 --steps 1 --warmup 1"
timeout -s KILL 200 $C > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:tc_slab_conv_kernel -s 0 -c 1 -o gpurun_out/${tag}_prof_conv -f $C > gpurun_out/${tag}_ncu1.log 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:tc_slab_wgrad_kernel -s 5 -c 1 -o gpurun_out/${tag}_prof_wgrad -f $C > gpurun_out/${tag}_ncu2.log 2>&1
exit 0
