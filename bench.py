#!/usr/bin/env python
"""bench.py -- T-bar inference + NMS throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic EM volume that is already resident in HBM
as uint8: FplNetwork.infer (tiled CNN forward, reference tile grid) -> voxel2obj (Gaussian
smoothing, 97th-percentile threshold, greedy NMS) -> detection list.  Workload at N=1 is
BASELINE.json configs[1]: vgg_like2 (scripts/fpl_cx1_0_vgg_4ss.py) on a synthetic 1024^3 volume,
obj_min_dist=27, smoothing_sigma=5, buffer_sz=15.  With N>1 every rank owns one such substack
(full_roi_inference semantics: independent substacks, flypylib/fplobjdetect.py:841-986) and the
detection lists are all-gathered over NCCL inside the timed region ("weak" scaling).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mvoxels/s T-bar inference+NMS"
UNIT = "Mvoxels/s"
DET = dict(obj_min_dist=27, smoothing_sigma=5, buffer_sz=15, thd=0)
NORM = (128.0, 33.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="volume edge (default: configs[1], 1024)")
    ap.add_argument("--model", default="vgg_like2", choices=["vgg_like", "vgg_like2", "unet_like2"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--tile-mult", type=int, default=None,
                    help="VGG only: evaluate super-tiles of this many reference tiles per axis "
                         "(bit-identical to the reference grid; default 4 for the VGGs, 1 for the U-Net)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops_sustained", 1400.0),
                    tf_burst=d.get("bf16_tflops", 1590.0), src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def synth_volume_device(size, seed, device):
    """EM-like uint8 volume generated on the device (not timed): smooth unit-variance noise -> 128+33*n."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((size, size, size), dtype=torch.uint8, device=device)
    slab = 64
    k = torch.ones((1, 1, 3, 3, 3), device=device) / 27.0
    for z0 in range(0, size, slab):
        z1 = min(size, z0 + slab)
        a = torch.randn((1, 1, z1 - z0 + 4, size + 4, size + 4), generator=g, device=device)
        a = torch.nn.functional.conv3d(torch.nn.functional.conv3d(a, k), k)
        a = a / a.std()
        out[z0:z1] = (128 + 33 * a[0, 0, :z1 - z0, :size, :size]).clamp_(0, 255).to(torch.uint8)
    return out


def seeded_weights(arch, seed=4321):
    """Random-init weights of the architecture in Keras get_weights() order: glorot_uniform kernels,
    non-trivial BN statistics (gamma~U(.5,1.5), beta,mean~N(0,.1), var~U(.5,1.5)), zero final bias."""
    from flypylib_b200 import fplmodels
    spec = fplmodels._ARCH[arch]
    rng = np.random.default_rng(seed)
    ws = []
    for k, cin, cout in spec["convs"]:
        lim = np.sqrt(6.0 / (k ** 3 * cin + k ** 3 * cout))
        ws.append(rng.uniform(-lim, lim, (k, k, k, cin, cout)).astype(np.float32))
        ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        ws.append((0.1 * rng.standard_normal(cout)).astype(np.float32))
        ws.append((0.1 * rng.standard_normal(cout)).astype(np.float32))
        ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    cin = spec["final_cin"]
    lim = np.sqrt(6.0 / (cin + 1))
    ws.append(rng.uniform(-lim, lim, (1, 1, 1, cin, 1)).astype(np.float32))
    if spec["final_bias"]:
        ws.append(np.zeros(1, np.float32))
    return ws


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def cpu_baseline(arch, n_tiles=24, v2o_edge=320):
    """Oracle (CPU port of the reference path) on a bounded sample of the same workload:
    n_tiles reference tiles through the torch-CPU restatement of the Keras graph (all host threads) +
    voxel2obj (C restatement, OpenMP Gaussian + sorted greedy) on a v2o_edge^3 crop."""
    import torch
    from oracle import models_oracle as M
    from oracle import voxel2obj_oracle as O
    from tests.golden import cases
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, rf, infer_sz, _ = M.ARCHS[arch]
    w = seeded_weights(arch)
    out_edge = infer_sz - 2 * rf[1]
    x = ((cases.em_volume((n_tiles * 8, infer_sz, infer_sz), seed=3).astype(np.float32) - NORM[0]) / NORM[1])
    x = np.stack([np.resize(x, (infer_sz, infer_sz, infer_sz)) for _ in range(n_tiles)])
    net = M.TorchNet(arch, w, dtype=torch.float32)
    t0 = time.perf_counter()
    net.predict(x[..., None], batch_size=1)
    t_fwd = time.perf_counter() - t0
    pm = cases.prob_map((v2o_edge,) * 3, 5, "blobs")
    t0 = time.perf_counter()
    O.voxel2obj(pm, DET["obj_min_dist"], DET["smoothing_sigma"], (0, 0, 0), DET["buffer_sz"], DET["thd"], impl="c")
    t_v2o = time.perf_counter() - t0
    per_voxel = t_fwd / (n_tiles * out_edge ** 3) + t_v2o / v2o_edge ** 3
    return {"value": 1e-6 / per_voxel, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d reference tiles (%d^3 -> %d^3) of %s on torch-CPU fp32 (%.1f s) + voxel2obj C oracle "
                      "on a %d^3 map (%.1f s)" % (n_tiles, infer_sz, out_edge, arch, t_fwd, v2o_edge, t_v2o)}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (Keras/TF are not installable
    here, so the arithmetic is the oracle port) on the box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    t_all0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue                        # one warm-up pass is enough for a CPU library path
        last = cpu_baseline(args.model, n_tiles=16, v2o_edge=288)
        if i >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals)) if vals else last["value"]
    ms = (args.size ** 3 / (v * 1e6)) * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s inference + voxel2obj on a synthetic %d^3 uint8 EM volume "
                                   "(ms_per_step extrapolated from the bounded sample)" % (args.model, args.size)},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from flypylib_b200 import fplmodels, fplnetwork, fplobjdetect, multi_gpu, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; flypylib_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):    # model.summary() (fplnetwork.py:58) must not precede the JSON line
        net = fplnetwork.FplNetwork(getattr(fplmodels, args.model))
    net.train_single.set_weights(seeded_weights(args.model))
    net.set_precision(args.precision)
    net._set_infer()
    if args.tile_mult is None:
        args.tile_mult = 1 if args.model == "unet_like2" else 4
    net.tile_multiplier = args.tile_mult
    ctx = _lib.context(local)

    size = args.size
    vol = synth_volume_device(size, 1234 + rank, dev)
    pred = torch.empty((size, size, size), dtype=torch.float32, device=dev)

    def step():
        net.infer_device(vol, normalize=NORM, out=pred)
        out = fplobjdetect.voxel2obj_device(pred, DET["obj_min_dist"], DET["smoothing_sigma"], (0, 0, 0),
                                            DET["buffer_sz"], DET["thd"])
        if world > 1:       # the only collective of the path: all-gather of the detection lists (NCCL)
            rows = torch.from_numpy(np.concatenate([out["locs"], out["conf"][:, None]], 1)).to(dev)
            parts = multi_gpu.allgather_detections(rows)
            return sum(int(p.shape[0]) for p in parts)
        return out["conf"].size

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        n_det = step()
    sync()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        n_det = step()
    e1.record()
    sync()
    elapsed_ms = e0.elapsed_time(e1)
    prof = ctx.profile_end()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t)
    ms_per_step = elapsed_ms / args.steps
    value = world * size ** 3 / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: pinned host uint8 volume -> H2D -> infer -> voxel2obj -> D2H detection list
    e2e = None
    if not args.no_e2e:
        host = torch.empty((size, size, size), dtype=torch.uint8).pin_memory()
        host.copy_(vol)
        dvol = torch.empty_like(vol)
        d2h = 0
        net.infer_host(host, normalize=NORM, out=pred, image_dev=dvol)      # untimed: sizes the chunk buffers
        sync()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        for _ in range(args.steps):
            # public API with a HOST volume: the H2D copy is pipelined behind the convolutions chunk by chunk
            net.infer_host(host, normalize=NORM, out=pred, image_dev=dvol)
            out = fplobjdetect.voxel2obj_device(pred, DET["obj_min_dist"], DET["smoothing_sigma"], (0, 0, 0),
                                                DET["buffer_sz"], DET["thd"])
            d2h = out["conf"].size * 32
        t1.record()
        sync()
        wall = (time.perf_counter() - w0) * 1e3
        e_ms = max(t0.elapsed_time(t1), wall) / args.steps
        if world > 1:
            t = torch.tensor([e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t)
        e2e = {"value": world * size ** 3 / (e_ms * 1e-3) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(size ** 3), "d2h_bytes_per_step": int(d2h)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms3, work3, cnt3 = prof["conv3"]
    achieved = (work3 / (ms3 * 1e-3) / 1e12) if ms3 > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv3_traffic.json")
    if os.path.exists(tpath):       # dram bytes per launch of the dominant kernel from the committed ncu capture
        traffic = json.load(open(tpath))
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf"],
                # DRAM bytes (read + write) of ONE launch of the dominant kernel in the committed ncu --set full
                # capture (profiles/, a 512^3 run: its launches are 1/8 of this run's); details alongside
                "traffic": (traffic or {}).get("dram_bytes_total"), "traffic_detail": traffic,
                "kernel": "conv_fused12_kernel + conv_umma_kernel<3> (tcgen05 implicit GEMM, 3x3x3 convolutions)",
                "launches": int(cnt3 / max(1, args.steps)), "ms_per_step": ms3 / args.steps,
                "peak_source": "%s bf16 sustained" % pk["src"]}
    families = {k: {"ms_per_step": v[0] / args.steps, "launch_groups": int(v[2] / max(1, args.steps))}
                for k, v in prof.items() if v[2]}
    g_ms, g_work, _ = prof["gauss"]
    if g_ms > 0:
        families["gauss"]["hbm_frac_algorithmic"] = (g_work / (g_ms * 1e-3) / 1e9) / pk["hbm"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "%s (%s, tcgen05 implicit GEMM) inference + voxel2obj(r=27, sigma=5, buffer=15) on a "
                                   "synthetic %d^3 uint8 EM volume per GPU, random-init weights, reference tile grid evaluated as %d^3-tile super-tiles"
                                   % (args.model, args.precision, size, args.tile_mult),
                       "l2": "inputs larger than L2 (1 GiB uint8 volume, 4 GiB probability map per step)",
                       "detections_per_step": int(n_det)},
            "roofline": roofline, "families": families, "gpu_launches": int(launches / max(1, args.steps)),
            "clocks": clocks, "e2e": e2e}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.model)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
