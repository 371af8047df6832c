#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/u_smoke.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_unet_sharded.py --size 2048 > gpurun_out/u_unet2048_n2.json 2> gpurun_out/u_unet2048_n2.err
exit 0
