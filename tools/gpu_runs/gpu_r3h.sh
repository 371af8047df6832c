#!/bin/bash
tag=${1:-r3h}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_network_gpu.py tests/test_umma_gpu.py -x -q -k "resnet or fp32_vs_float64" > gpurun_out/${tag}_resnet.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_resnet.log
timeout -s KILL 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
exit 0
