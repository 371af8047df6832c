#!/bin/bash
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
exit 0
