#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_global_v2o.py --size 512 > gpurun_out/s_global_n2.json 2> gpurun_out/s_global_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/check_global_v2o.py --size 640 --kind uniform > gpurun_out/s_global_n2u.json 2> gpurun_out/s_global_n2u.err
exit 0
