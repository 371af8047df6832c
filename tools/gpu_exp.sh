#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/z3_bench_n$N.err | tail -1 > gpurun_out/z3_bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/bench_unet_sharded.py --size 2048 2> gpurun_out/z3_unet_n$N.err | tail -1 > gpurun_out/z3_unet_n$N.json
exit 0
