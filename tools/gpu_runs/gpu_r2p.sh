#!/bin/bash
# N-GPU strong-scaling bench (the driver's SCALE run does the same at round end)
tag=${1:-r2p}; n=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err; echo "rc=$?" >> gpurun_out/${tag}_bench_n${n}.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 tools/bench_train.py > gpurun_out/${tag}_train_n${n}.json 2> gpurun_out/${tag}_train_n${n}.err
exit 0
