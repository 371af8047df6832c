#!/bin/bash
# round-2 GPU pass B (2 GPUs): sharded strong-scaling bench at N=2 (+ cfg3 leg), N=1 line for the detection checksum
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; echo "rc=$?" >> gpurun_out/${tag}_bench_n2.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "rc=$?" >> gpurun_out/${tag}_bench_n1.err
exit 0
