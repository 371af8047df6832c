#!/bin/bash
tag=${1:-r2E}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxel2obj_gpu.py -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 300 python bench.py --size 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${tag}_bench512.json 2> gpurun_out/${tag}_bench512.err
timeout 300 python tools/bench_voxel2obj.py --size 1024 --kind uniform --steps 3 > gpurun_out/${tag}_v2o_1024u.json 2> gpurun_out/${tag}_v2o_1024u.err
exit 0
