#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_voxel2obj_gpu.py tests/test_global_v2o_gpu.py -x -q > gpurun_out/a3_tests.log 2>&1
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 > gpurun_out/a3_v2o_1024.json 2> gpurun_out/a3.err
timeout 600 python tools/bench_voxel2obj.py --size 1024 --steps 3 --kind uniform > gpurun_out/a3_v2o_1024u.json 2> gpurun_out/a3u.err
exit 0
