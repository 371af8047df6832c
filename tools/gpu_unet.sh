#!/bin/bash
# U-Net single-GPU rate (bench.py --model unet_like2 --size 512) + the network parity tests
tag=${1:-unet}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_umma_gpu.py tests/test_network_gpu.py tests/test_sharded_gpu.py -q > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_tests.log
timeout -s KILL 600 python bench.py --model unet_like2 --size 512 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${tag}_unet512.json 2> gpurun_out/${tag}_unet512.err; echo "rc=$?" >> gpurun_out/${tag}_unet512.err
exit 0
