#!/usr/bin/env python
"""Detection parity AT SCALE: fpl_voxel2obj on the GPU against the C oracle (oracle/voxel2obj_c.c, OpenMP Gaussian +
sorted greedy, itself pinned bit-for-bit to goldens of the unmodified reference) on maps of benchmark size --
the 1024^3 bench map and a map with more than 2^32 voxels (uint64 flat indices, the radix-class superset at > 10^8
candidates).  Compares threshold, detection list, order and confidences bit for bit.

    python tools/check_v2o_scale.py --shape 1024 1024 1024 --kind blobs --out gpurun_out/v2o_scale_1024.json
    python tools/check_v2o_scale.py --shape 4100 1024 1024 --kind blobs --out gpurun_out/v2o_scale_gt2p32.json

The oracle is the checker here (test infrastructure), never the thing measured.  Host memory: ~ 2 x 4 B per padded
voxel (the probability map is dropped before the percentile copy is made).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_map(shape, seed, kind, dev):
    """float32 (Z,Y,X) map on the device, z-block by z-block (same recipe as tools/bench_voxel2obj.synth_map)."""
    import torch
    Z, Y, X = shape
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    block = 64
    for z0 in range(0, Z, block):
        z1 = min(Z, z0 + block)
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 100003 + z0 // block)
        if kind == "uniform":
            out[z0:z1] = torch.rand((z1 - z0, Y, X), generator=g, device=dev)
            continue
        a = torch.rand((z1 - z0, Y, X), generator=g, device=dev) * 0.02
        n = max(1, int(3.0 * (z1 - z0) * Y * X / 50.0 ** 3))
        idx = torch.randint(0, (z1 - z0) * Y * X, (n,), generator=g, device=dev)
        amp = 0.85 + 0.3 * torch.rand(n, generator=g, device=dev)
        seeds = torch.zeros((z1 - z0) * Y * X, device=dev)
        seeds.index_put_((idx,), amp, accumulate=True)
        seeds = seeds.view(1, 1, z1 - z0, Y, X)
        ax = torch.arange(-9, 10, device=dev, dtype=torch.float32)
        k1 = torch.exp(-0.5 * (ax / 3.0) ** 2)
        for dim in range(3):
            kshape = [1, 1, 1, 1, 1]; kshape[2 + dim] = 19
            pad = [0, 0, 0, 0, 0, 0]; pad[(2 - dim) * 2] = 9; pad[(2 - dim) * 2 + 1] = 9
            seeds = torch.nn.functional.conv3d(torch.nn.functional.pad(seeds, pad), k1.view(kshape))
        out[z0:z1] = (a + seeds[0, 0]).clamp_(0, 1)
        del seeds, a
    return out


def compare(shape, kind, seed, r, sigma, buf, thd, classic=False, mode=None):
    """mode: 0 default (two-tier when the map qualifies), 1 classic exact, 2 fused exact"""
    import ctypes
    import torch
    from flypylib_b200 import fplobjdetect, _lib
    from oracle import voxel2obj_oracle as O
    dev = torch.device("cuda", 0)
    t0 = time.perf_counter()
    pm = synth_map(shape, seed, kind, dev)
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.fpl_debug_v2o_classic.argtypes = [ctypes.c_int]
    mode = (1 if classic else 0) if mode is None else int(mode)
    lib.fpl_debug_v2o_classic(mode)
    lib.fpl_debug_v2o_decline_reason.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.fpl_debug_v2o_decline_reason(_lib.context(0).handle, 1)          # no back-off inherited from earlier maps
    try:
        got, st = fplobjdetect.voxel2obj_device(pm, r, sigma, (0, 0, 0), buf, thd, return_stats=True)    # warm-up
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        got, st = fplobjdetect.voxel2obj_device(pm, r, sigma, (0, 0, 0), buf, thd, return_stats=True)
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t1
    finally:
        lib.fpl_debug_v2o_classic(0)
    host = pm.cpu().numpy()
    del pm
    torch.cuda.empty_cache()
    _lib.context(0).release_workspace()
    t2 = time.perf_counter()
    s = O.smooth_padded_c(host, r, sigma)
    del host
    t3 = time.perf_counter()
    t = O.threshold(s, thd)
    t4 = time.perf_counter()
    rows = O.greedy_nms_c(s, t, r, max_out=max(1 << 20, 4 * got["conf"].size + 1024))
    t5 = time.perf_counter()
    want = O.finish(rows, r, shape, buf, (0, 0, 0))
    n_vox = int(np.prod([int(v) for v in shape]))
    n_pad = int(np.prod([int(v) + 2 * r for v in shape]))
    same_thresh = bool(float(t) == st["threshold"])
    same_locs = bool(got["locs"].shape == want["locs"].shape and np.array_equal(got["locs"], want["locs"]))
    same_conf = bool(got["conf"].shape == want["conf"].shape and np.array_equal(got["conf"], want["conf"]))
    n_cand = int(np.count_nonzero(s > t))
    return {"shape": list(shape), "kind": kind, "seed": seed, "r": r, "sigma": sigma, "buffer": buf, "thd": thd,
            "gpu_path": st.get("path"), "voxels": n_vox, "padded_voxels": n_pad,
            "voxels_over_2p32": n_vox > 2 ** 32, "threshold_gpu": st["threshold"], "threshold_oracle": float(t),
            "detections_gpu": int(got["conf"].size), "detections_oracle": int(want["conf"].size),
            "candidates_oracle": n_cand, "gpu_stats": {k: (v if isinstance(v, (float, str)) or v is None else int(v)) for k, v in st.items()},
            "max_flat_index_gpu": int((got["locs"][:, 2] * shape[1] * shape[2] + got["locs"][:, 1] * shape[2]
                                       + got["locs"][:, 0]).max()) if got["conf"].size else -1,
            "threshold_identical": same_thresh, "locs_identical": same_locs, "conf_identical": same_conf,
            "identical": same_thresh and same_locs and same_conf,
            "seconds": {"synth": t1 - t0, "gpu_voxel2obj": t_gpu, "oracle_smooth": t3 - t2, "oracle_percentile": t4 - t3,
                        "oracle_greedy": t5 - t4}, "host_cores": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[1024, 1024, 1024])
    ap.add_argument("--kind", default="blobs")
    ap.add_argument("--seed", type=int, default=99)
    ap.add_argument("--r", type=int, default=27)
    ap.add_argument("--sigma", type=float, default=5.0)
    ap.add_argument("--buffer", type=int, default=15)
    ap.add_argument("--thd", type=float, default=0.0)
    ap.add_argument("--classic", action="store_true", help="force the classic (five dense passes) GPU path")
    ap.add_argument("--mode", type=int, default=None, help="0 default (two-tier), 1 classic exact, 2 fused exact")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = compare(tuple(a.shape), a.kind, a.seed, a.r, a.sigma, a.buffer, a.thd, classic=a.classic, mode=a.mode)
    line = json.dumps(res)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(line + "\n")
    sys.exit(0 if res["identical"] else 1)


if __name__ == "__main__":
    main()
