#!/bin/bash
# round-2 (session 2): tensor-core training step -- parity tests, bench at the three precisions, launch list
tag=${1:-r3a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -x -q > gpurun_out/${tag}_train_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_train_tests.log
for p in tf32 bf16 fp32; do
  timeout 300 python tools/bench_train.py --precision $p > gpurun_out/${tag}_train_$p.json 2> gpurun_out/${tag}_train_$p.err; echo "rc=$?" >> gpurun_out/${tag}_train_$p.err
done
timeout 300 python tools/bench_train.py --precision tf32 --steps 2 --warmup 1 > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv \
    python tools/bench_train.py --precision tf32 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu.log 2>&1
exit 0
