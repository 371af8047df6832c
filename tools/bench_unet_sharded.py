#!/usr/bin/env python
"""BASELINE config 3: unet_like2 inference on a synthetic D^3 volume, z-slab sharded with buffer halos
over the ranks of one box (one process per GPU), detections all-gathered over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/bench_unet_sharded.py --size 2048
Every rank synthesises only its own slab (+ buffer) of the volume.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--model", default="unet_like2")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--buffer", type=int, default=35)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from flypylib_b200 import fplmodels, fplnetwork, multi_gpu
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        net = fplnetwork.FplNetwork(getattr(fplmodels, a.model))
    net.train_single.set_weights(bench.seeded_weights(a.model))
    net.set_precision("bf16")
    net._set_infer()
    if a.model != "unet_like2":
        net.tile_multiplier = 4
    Z = a.size
    z0, z1 = multi_gpu.partition_layers(Z, world)[rank]
    lo, hi = max(0, z0 - a.buffer), min(Z, z1 + a.buffer)
    # slab of the synthetic volume: generated as a (hi-lo, D, D) volume seeded by rank
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    sub = torch.empty((hi - lo, a.size, a.size), dtype=torch.uint8, device=dev)
    k = torch.ones((1, 1, 3, 3, 3), device=dev) / 27.0
    for s0 in range(0, hi - lo, 32):
        s1 = min(hi - lo, s0 + 32)
        x = torch.randn((1, 1, s1 - s0 + 4, a.size + 4, a.size + 4), generator=g, device=dev)
        x = torch.nn.functional.conv3d(torch.nn.functional.conv3d(x, k), k)
        sub[s0:s1] = (128 + 33 * x[0, 0] / x.std()).clamp_(0, 255).to(torch.uint8)
        del x
    from flypylib_b200 import fplobjdetect

    def step():
        pred = net.infer_device(sub, normalize=bench.NORM)
        out = fplobjdetect.voxel2obj_device(pred, 27, 5, (0, 0, lo), 0, 0)
        rows = np.concatenate([out["locs"], out["conf"][:, None]], 1)
        rows = rows[(rows[:, 2] >= z0) & (rows[:, 2] < z1)]
        if world > 1:
            parts = multi_gpu.allgather_detections(torch.from_numpy(np.ascontiguousarray(rows)).to(dev))
            return sum(int(p.shape[0]) for p in parts)
        return rows.shape[0]

    n = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        n = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    if world > 1:
        t = torch.tensor([dt], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t)
    if rank == 0:
        print(json.dumps({"metric": "Mvoxels/s %s inference+NMS, z-slab sharded (config 3)" % a.model,
                          "value": a.size ** 3 / dt / 1e6, "unit": "Mvoxels/s", "n_gpus": world, "size": a.size,
                          "s_per_step": dt, "detections": int(n), "buffer": a.buffer,
                          "slab_planes_rank0": int(hi - lo)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
