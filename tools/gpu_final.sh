#!/bin/bash
# round-end evidence pass on one B200: all GPU tests, smoke, bench N=1 (all legs), reference arm, training bench + launch list
tag=${1:-fin}
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "rc=$?" >> gpurun_out/${tag}_bench_ref.err
for p in tf32 bf16; do
  timeout -s KILL 200 python tools/bench_train.py --precision $p --steps 20 --warmup 5 > gpurun_out/${tag}_train_$p.json 2> gpurun_out/${tag}_train_$p.err
done
timeout -s KILL 200 python tools/bench_train.py --precision tf32 --steps 2 --warmup 1 > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_train_launches.csv \
    python tools/bench_train.py --precision tf32 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu.log 2>&1
exit 0
