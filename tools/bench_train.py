#!/usr/bin/env python
"""BASELINE config 5: one data-parallel training step of a VGG builder on synthetic EM minibatches, one process per
GPU, gradient all-reduce over NCCL (replaces flypylib/multi_gpu.py:20-61 tower replication + Keras fit_generator,
flypylib/fplnetwork.py:112-128).  64 patches per GPU (scripts/fpl_cx1_0_vgg_4ss.py:11-17).

    python tools/bench_train.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py

Prints one JSON line on rank 0: ms/step (device events, max over ranks), the share of forward+backward, gradient
all-reduce and Adam, patches/s, and whether all ranks hold identical parameters afterwards.  The per-GPU minibatch is
already resident in HBM (the reference's generator thread is out of scope); "e2e" adds the H2D copy of a pinned
host minibatch per step.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vgg_like2", choices=["vgg_like", "vgg_like2"])
    ap.add_argument("--batch", type=int, default=64, help="patches per GPU")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "bf16", "fp32"],
                    help="tf32 = tcgen05 on bf16 hi/lo split operands (fp32-class), bf16 = one contraction, fp32 = CUDA cores")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from flypylib_b200 import fplmodels, fpltrain
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    builder = getattr(fplmodels, a.model)
    rf = builder()[1][0]
    model = builder(rf)[0]
    model.set_weights(bench.seeded_weights(a.model))
    tr = fpltrain.Trainer(model, rf, a.batch, precision=a.precision)            # broadcasts rank 0's parameters
    g = torch.Generator(device=dev); g.manual_seed(77 + rank)
    x = torch.randn((a.batch, rf, rf, rf), generator=g, device=dev)
    y = (torch.rand(a.batch, generator=g, device=dev) < 0.5).to(torch.uint8)
    hx = x.cpu().pin_memory(); hy = y.cpu().pin_memory()
    gb = a.batch * world
    ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, marks=None, from_host=False):
        if from_host:
            x.copy_(hx, non_blocking=True); y.copy_(hy, non_blocking=True)
        if marks: marks[0].record()
        tr.forward_backward(x, y, gb, 1000 + i)
        if marks: marks[1].record()
        tr.allreduce()
        if marks: marks[2].record()
        tr.apply()
        if marks: marks[3].record()

    for i in range(a.warmup):
        step(i)
    sync()
    t0, t1 = ev(), ev()
    marks = [[ev() for _ in range(4)] for _ in range(a.steps)]
    t0.record()
    for i in range(a.steps):
        step(a.warmup + i, marks[i])
    t1.record()
    sync()
    ms = t0.elapsed_time(t1) / a.steps
    fb = float(np.mean([m[0].elapsed_time(m[1]) for m in marks]))
    ar = float(np.mean([m[1].elapsed_time(m[2]) for m in marks]))
    ad = float(np.mean([m[2].elapsed_time(m[3]) for m in marks]))
    sync()
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(a.steps):
        step(a.warmup + a.steps + i, None, from_host=True)
    e1.record()
    sync()
    e_ms = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms, fb, ar, ad, e_ms], device=dev)
    same = True
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ref = tr.params.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([1 if torch.equal(ref, tr.params) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    ms, fb, ar, ad, e_ms = (float(v) for v in t)
    if rank == 0:
        spec = fplmodels._ARCH[a.model]
        n_params = int(sum(int(np.prod(s)) for s in model.weight_shapes()))
        # algorithmic work of one step: forward + dgrad + wgrad = 3 x 2*k^3*Cin*Cout*out^3 per convolution and patch
        # (max-pooling after the 2nd and 4th convolution of both VGG builders, flypylib/fplmodels.py:102-172)
        d, fl = rf, 0.0
        for ci, (k, cin, cout) in enumerate(spec["convs"]):
            d -= k - 1
            fl += 2.0 * k ** 3 * cin * cout * d ** 3
            if ci in (1, 3):
                d //= 2
        flops_step = 3.0 * fl * gb
        print(json.dumps({"metric": "ms per data-parallel training step (config 5)", "model": a.model, "n_gpus": world,
                          "patches_per_gpu": a.batch, "patch": rf, "ms_per_step": ms, "patches_per_s": gb / (ms * 1e-3),
                          "ms_forward_backward": fb, "ms_grad_allreduce": ar, "ms_adam_bn_update": ad,
                          "allreduce_bytes": 4 * n_params, "e2e_ms_per_step_with_h2d": e_ms,
                          "identical_parameters_on_all_ranks": same, "arithmetic": {"tf32": "tcgen05 bf16 hi/lo x3, fp32 accumulate (csrc/train_tc.cuh)",
                                         "bf16": "tcgen05 bf16, fp32 accumulate (csrc/train_tc.cuh)",
                                         "fp32": "fp32 CUDA-core kernels (csrc/train.cu)"}[a.precision],
                          "steps": a.steps, "warmup": a.warmup, "convs": len(spec["convs"]),
                          "algorithmic_gflop_per_step": flops_step / 1e9,
                          "algorithmic_tflops": flops_step / (ms * 1e-3) / 1e12}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
