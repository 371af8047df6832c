#!/bin/bash
# round-2 GPU pass A: parity tests, bench N=1 (with extra legs), detection parity at scale
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/${tag}_env.log 2>&1; free -g >> gpurun_out/${tag}_env.log; nproc >> gpurun_out/${tag}_env.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 600 python tools/check_v2o_scale.py --shape 1024 1024 1024 --kind blobs --out gpurun_out/${tag}_v2o_scale_1024.json > gpurun_out/${tag}_scale1024.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_scale1024.log
timeout 1500 python tools/check_v2o_scale.py --shape 4100 1024 1024 --kind blobs --out gpurun_out/${tag}_v2o_scale_gt2p32.json > gpurun_out/${tag}_scale_gt2p32.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_scale_gt2p32.log
exit 0
