#!/bin/bash
tag=${1:-r3c}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_train_tc_gpu.py -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/${tag}_tc.log
FPL_TC_GATHER=1 timeout -s KILL 300 python -m pytest tests/test_train_tc_gpu.py -q 2>&1 | tail -3 > gpurun_out/${tag}_tc_gather.log
timeout -s KILL 300 python -m pytest tests/test_train_gpu.py -q -s 2>&1 > gpurun_out/${tag}_train.log
for p in tf32 bf16; do
  timeout -s KILL 200 python tools/bench_train.py --precision $p > gpurun_out/${tag}_train_$p.json 2> gpurun_out/${tag}_train_$p.err; echo "rc=$?" >> gpurun_out/${tag}_train_$p.err
done
timeout -s KILL 200 python tools/bench_train.py --precision tf32 --steps 2 --warmup 1 > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv \
    python tools/bench_train.py --precision tf32 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu.log 2>&1
exit 0
