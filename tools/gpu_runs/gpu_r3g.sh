#!/bin/bash
# full GPU evidence pass (session 2 of round 2): all GPU tests, smoke, bench N=1, reference arm, cfg5 training bench
tag=${1:-r3g}
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_voxel2obj_seg_gpu.py -q > gpurun_out/${tag}_seg.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_seg.log
timeout -s KILL 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
exit 0
