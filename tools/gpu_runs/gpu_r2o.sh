#!/bin/bash
tag=${1:-r2o}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sharded_gpu.py tests/test_umma_gpu.py tests/test_network_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
exit 0
