#!/bin/bash
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k "row_sharded or sharded_path" > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; echo "rc=$?" >> gpurun_out/${tag}_bench_n2.err
exit 0
