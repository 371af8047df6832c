#!/bin/bash
tag=${1:-r2j}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 900 python -m pytest tests/test_voxel2obj_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
exit 0
