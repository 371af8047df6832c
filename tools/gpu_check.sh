#!/bin/bash
# One GPU-box pass: parity tests, bench, launch list, ncu full capture of the dominant kernel.
# Usage (from the repo root, on the box): bash tools/gpu_check.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain512.log 2>&1 || exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_512.csv \
    python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fused12_kernel -s 1 -c 1 -o gpurun_out/${tag}_prof_fused12 -f \
    python bench.py --size 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu_fused.log 2>&1
ncu -i gpurun_out/${tag}_prof_fused12.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_fused12_raw.csv 2>/dev/null
exit 0
