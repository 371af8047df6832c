#!/bin/bash
# round-2 evidence pass: full GPU tests, bench N=1 (all legs), reference arm, detection parity at scale, ncu of the new kernels
tag=${1:-r2t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "rc=$?" >> gpurun_out/${tag}_bench_ref.err
timeout 600 python tools/check_v2o_scale.py --shape 1024 1024 1024 --kind blobs --out gpurun_out/${tag}_v2o_scale_1024.json > gpurun_out/${tag}_scale1024.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_scale1024.log
timeout 1500 python tools/check_v2o_scale.py --shape 4400 1024 1024 --kind blobs --out gpurun_out/${tag}_v2o_scale_gt2p32.json > gpurun_out/${tag}_scale_gt2p32.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_scale_gt2p32.log
timeout 300 python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/${tag}_plain.log 2>&1 || exit 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gauss32_kernel|approx_pass1_kernel' -c 2 -o gpurun_out/${tag}_prof_v2o -f \
    python tools/bench_voxel2obj.py --size 1024 --steps 1 --warmup 0 > gpurun_out/${tag}_ncu2.log 2>&1
exit 0
