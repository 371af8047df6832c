#!/usr/bin/env python
"""BASELINE config 4: standalone voxel2obj (Gaussian + percentile threshold + greedy NMS) on a
precomputed float32 probability map resident in HBM.  Prints one JSON line with Mvoxels/s and the
fraction of the HBM roofline (12 B per voxel: read map, write smoothed, read smoothed).

    python tools/bench_voxel2obj.py [--size 2048] [--kind blobs|uniform] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_map(size, seed, kind, dev):
    """float32 map generated on the device slab by slab: 'uniform' i.i.d. U(0,1) (adversarial) or 'blobs'
    (sparse Gaussian bumps of sigma 3, peak 0.85..1, clipped at 1, + U(0,0.02) noise)."""
    import torch
    g = torch.Generator(device=dev); g.manual_seed(seed)
    out = torch.empty((size, size, size), dtype=torch.float32, device=dev)
    slab = 64
    for z0 in range(0, size, slab):
        z1 = min(size, z0 + slab)
        if kind == "uniform":
            out[z0:z1] = torch.rand((z1 - z0, size, size), generator=g, device=dev)
            continue
        a = torch.rand((z1 - z0, size, size), generator=g, device=dev) * 0.02
        n = max(1, int(3.0 * (z1 - z0) * size * size / 50.0 ** 3))
        idx = torch.randint(0, (z1 - z0) * size * size, (n,), generator=g, device=dev)
        amp = 0.85 + 0.3 * torch.rand(n, generator=g, device=dev)
        seeds = torch.zeros((z1 - z0) * size * size, device=dev)
        seeds.index_put_((idx,), amp, accumulate=True)
        seeds = seeds.view(1, 1, z1 - z0, size, size)
        ax = torch.arange(-9, 10, device=dev, dtype=torch.float32)
        k1 = torch.exp(-0.5 * (ax / 3.0) ** 2)
        for dim in range(3):
            shape = [1, 1, 1, 1, 1]; shape[2 + dim] = 19
            pad = [0, 0, 0, 0, 0, 0]; pad[(2 - dim) * 2] = 9; pad[(2 - dim) * 2 + 1] = 9
            seeds = torch.nn.functional.conv3d(torch.nn.functional.pad(seeds, pad), k1.view(shape))
        out[z0:z1] = (a + seeds[0, 0]).clamp_(0, 1)
        del seeds
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--kind", default="blobs")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    import torch
    from flypylib_b200 import fplobjdetect, _lib
    dev = torch.device("cuda", 0)
    pm = synth_map(a.size, 99, a.kind, dev)
    ctx = _lib.context(0)
    for _ in range(a.warmup):
        out, st = fplobjdetect.voxel2obj_device(pm, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
    torch.cuda.synchronize()
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out, st = fplobjdetect.voxel2obj_device(pm, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    prof = ctx.profile_end()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    vox = a.size ** 3
    achieved = 12.0 * vox / (ms * 1e-3) / 1e9
    print(json.dumps({"metric": "Mvoxels/s voxel2obj (config 4)", "value": vox / (ms * 1e-3) / 1e6, "unit": "Mvoxels/s",
                      "ms_per_step": ms, "size": a.size, "kind": a.kind, "detections": int(out["conf"].size),
                      "stats": st, "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                "frac": achieved / peaks["hbm_gbs"]},
                      "families": {k: round(v[0] / a.steps, 2) for k, v in prof.items() if v[2]},
                      "workspace_GiB": ctx.workspace_bytes() / 2 ** 30}))


if __name__ == "__main__":
    main()
