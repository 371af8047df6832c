#!/usr/bin/env python
"""Exact multi-GPU voxel2obj (semantics S2) on real ranks: every rank holds a z-slab of the same synthetic
probability map, multi_gpu.voxel2obj_global runs over NCCL, rank 0 also runs the single-GPU voxel2obj on the
whole map and compares bit for bit.  Prints one JSON line (rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/check_global_v2o.py --size 512 [--kind blobs|uniform]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--kind", default="blobs")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench_voxel2obj import synth_map
    from flypylib_b200 import fplobjdetect, multi_gpu
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pm = synth_map(a.size, 99, a.kind, dev)                  # same seed on every rank -> same map
    z0, z1 = multi_gpu.partition_layers(a.size, world)[rank]
    slab = pm[z0:z1].contiguous()
    if rank != 0:
        del pm
    out = None
    for it in range(2):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, st = multi_gpu.voxel2obj_global([slab], [(z0, z1)], a.size, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
        dist.barrier(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if rank == 0:
        t0 = time.perf_counter()
        want, st1 = fplobjdetect.voxel2obj_device(pm, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
        torch.cuda.synchronize()
        dt1 = time.perf_counter() - t0
        same = bool(np.array_equal(out["locs"], want["locs"]) and np.array_equal(out["conf"], want["conf"]) and
                    st["threshold"] == st1["threshold"])
        print(json.dumps({"check": "voxel2obj_global == single-GPU voxel2obj", "identical": same, "n_gpus": world,
                          "size": a.size, "kind": a.kind, "detections": int(want["conf"].size), "rounds": st["rounds"],
                          "s_global": dt, "s_single": dt1}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
