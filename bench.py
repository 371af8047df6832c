#!/usr/bin/env python
"""bench.py -- T-bar inference + NMS throughput (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over ONE synthetic EM volume (BASELINE.json configs[1]: vgg_like2,
scripts/fpl_cx1_0_vgg_4ss.py, 1024^3 uint8, obj_min_dist=27, smoothing_sigma=5, buffer_sz=15): FplNetwork.infer (tiled CNN
forward, reference tile grid) -> voxel2obj (Gaussian smoothing, 97th-percentile threshold, greedy NMS) -> detection list.
N = 1: the volume is resident in HBM.  N > 1: the SAME volume is z-slab sharded over the ranks (strong scaling): every
rank evaluates its slab + 2*rf_offset halo with no forward communication, then the exact-global voxel2obj
(multi_gpu.voxel2obj_global) yields the single-GPU detection list on every rank -- `detections_sha256_16` is the same at
every N.  Extra legs, each guarded so that it can never take the main line down: `e2e_dropin` (the literal reference call
sequence with numpy arrays), `cfg1`, `cfg4` and `cfg5` at N = 1 (BASELINE configs[0], [3], [4]; `cfg5` also at N > 1), `cfg3` at N > 1 (configs[2]: the
U-Net on a 2048^3 volume, tile rows dealt over the ranks).

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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra legs (cfg1 / cfg4 at N=1, cfg3 at N>1, e2e_dropin)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf=d.get("bf16_tflops_sustained", 1400.0),
                    tf_burst=d.get("bf16_tflops", 1590.0), src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def synth_volume_planes(size, z_lo, z_hi, seed, device, block=64):
    """Planes [z_lo, z_hi) of THE synthetic EM-like uint8 volume (not timed), generated on the device.  The volume is
    defined block by block (64 planes, seed + block index), so any rank can produce exactly its own planes of the one
    global volume: smooth unit-variance noise -> 128 + 33 * n."""
    import torch
    out = torch.empty((z_hi - z_lo, size, size), dtype=torch.uint8, device=device)
    k = torch.ones((1, 1, 3, 3, 3), device=device) / 27.0
    for bi in range(z_lo // block, -(-z_hi // block)):
        b0, b1 = bi * block, min(size, (bi + 1) * block)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 100003 + bi)
        a = torch.randn((1, 1, b1 - b0 + 4, size + 4, size + 4), generator=g, device=device)
        a = torch.nn.functional.conv3d(torch.nn.functional.conv3d(a, k), k)
        a = a / a.std()
        blk = (128 + 33 * a[0, 0, :b1 - b0, :size, :size]).clamp_(0, 255).to(torch.uint8)
        lo, hi = max(b0, z_lo), min(b1, z_hi)
        out[lo - z_lo:hi - z_lo] = blk[lo - b0:hi - b0]
        del a, blk
    return out


def synth_volume_device(size, seed, device):
    return synth_volume_planes(size, 0, size, seed, device)


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


def cpu_cfg1():
    """BASELINE configs[0] in full, not extrapolated: vgg_like on a 256^3 volume through the reference tiler
    (27 tiles of 102^3, torch-CPU fp32 restatement of the Keras graph) + voxel2obj (C restatement) on the whole map."""
    import torch
    from oracle import models_oracle as M
    from oracle import voxel2obj_oracle as O
    from tests.golden import cases
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, rf, infer_sz, _ = M.ARCHS["vgg_like"]
    img = ((cases.em_volume((256, 256, 256), seed=1234).astype(np.float32) - NORM[0]) / NORM[1]).astype(np.float32)
    net = M.TorchNet("vgg_like", seeded_weights("vgg_like"), dtype=torch.float32)
    t0 = time.perf_counter()
    pred = M.infer_tiler(img, net, (infer_sz,) * 3, (rf[1],) * 3, n_gpu=1)
    t1 = time.perf_counter()
    out = O.voxel2obj(pred, DET["obj_min_dist"], DET["smoothing_sigma"], (0, 0, 0), DET["buffer_sz"], DET["thd"], impl="c")
    t2 = time.perf_counter()
    return {"workload": "vgg_like inference + voxel2obj on a synthetic 256^3 volume (BASELINE configs[0]), whole "
                        "workload, not extrapolated", "value": 256 ** 3 / (t2 - t0) / 1e6, "unit": UNIT,
            "s_infer": t1 - t0, "s_voxel2obj": t2 - t1, "cores": cores, "kind": "port",
            "detections": int(out["conf"].size)}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (Keras/TF are not installable
    here, so the arithmetic is the oracle port) on the box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s inference + voxel2obj on a synthetic %d^3 uint8 EM volume "
                                   "(ms_per_step extrapolated from the bounded sample)" % (args.model, args.size)},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_extras:
        try:
            line["cfg1"] = cpu_cfg1()
        except Exception as e:      # noqa: BLE001 -- an extra leg must never take the main line down
            line["cfg1"] = {"error": repr(e)[:200]}
    line["wall_s"] = time.perf_counter() - t_all0
    print(json.dumps(line))


def build_net(model, precision, tile_mult):
    import contextlib
    from flypylib_b200 import fplmodels, fplnetwork
    with contextlib.redirect_stdout(sys.stderr):    # model.summary() (fplnetwork.py:58) must not precede the JSON line
        net = fplnetwork.FplNetwork(getattr(fplmodels, model))
    net.train_single.set_weights(seeded_weights(model))
    net.set_precision(precision)
    net._set_infer()
    net.tile_multiplier = tile_mult if tile_mult is not None else (1 if model.startswith("unet") else 4)
    return net


def timed(fn, steps, sync):
    """steps calls of fn between two CUDA events on the current stream, device synchronised on both sides."""
    import torch
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    r = None
    for _ in range(steps):
        r = fn()
    e1.record()
    sync()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - w0) * 1e3 / steps, r


def leg_cfg1(dev, sync):
    """BASELINE configs[0] on the GPU: vgg_like + voxel2obj on the 256^3 volume the CPU arm runs in full."""
    import torch
    from flypylib_b200 import fplobjdetect
    from tests.golden import cases
    net = build_net("vgg_like", "bf16", 4)
    host = torch.from_numpy(cases.em_volume((256, 256, 256), seed=1234)).pin_memory()
    vol = torch.empty_like(host, device=dev)

    def step():
        vol.copy_(host, non_blocking=True)
        pred = net.infer_device(vol, normalize=NORM)
        return fplobjdetect.voxel2obj_device(pred, DET["obj_min_dist"], DET["smoothing_sigma"], (0, 0, 0),
                                             DET["buffer_sz"], DET["thd"])
    for _ in range(3):
        step()
    ms, wall, out = timed(step, 5, sync)
    ms = max(ms, wall)
    return {"workload": "vgg_like (bf16) inference + voxel2obj on a synthetic 256^3 volume (BASELINE configs[0]), "
                        "pinned host uint8 -> H2D -> infer -> voxel2obj -> D2H list", "value": 256 ** 3 / (ms * 1e-3) / 1e6,
            "unit": UNIT, "ms_per_step": ms, "detections": int(out["conf"].size)}


def leg_cfg4(dev, sync, ctx, pk, size=2048):
    """BASELINE configs[3]: standalone voxel2obj on a precomputed size^3 float32 probability map in HBM."""
    import torch
    from flypylib_b200 import fplobjdetect
    from tools import bench_voxel2obj
    pm = bench_voxel2obj.synth_map(size, 99, "blobs", dev)

    def step():
        return fplobjdetect.voxel2obj_device(pm, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
    for _ in range(2):
        step()
    ctx.profile_begin()
    ms, _, (out, st) = timed(step, 3, sync)
    prof = ctx.profile_end()
    achieved = 12.0 * size ** 3 / (ms * 1e-3) / 1e9
    return {"workload": "voxel2obj(r=27, sigma=5, buffer=30) on a synthetic %d^3 float32 probability map resident in HBM "
                        "(BASELINE configs[3])" % size, "value": size ** 3 / (ms * 1e-3) / 1e6, "unit": UNIT,
            "ms_per_step": ms, "detections": int(out["conf"].size), "path": st.get("path"),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": achieved / pk["hbm"], "algorithmic_bytes_per_voxel": 12},
            "families": {k: round(v[0] / 3, 3) for k, v in prof.items() if v[2]}}


def leg_cfg3(dev, sync, world, rank, size=2048):
    """BASELINE configs[2]: unet_like2 on ONE synthetic size^3 volume, sharded over the ranks with the (z, y) tile rows
    of the reference grid dealt evenly (the U-Net can only be cut on the tile grid; rows a rank evaluates for a
    neighbour's layer go to the plane owner over NCCL P2P), exact-global voxel2obj, detections on every rank."""
    import torch
    from flypylib_b200 import multi_gpu
    net = build_net("unet_like2", "bf16", 1)
    off, out = int(net.rf_offset[0]), net.slab_granularity()
    pieces, plans = multi_gpu.row_plan(size, size, off, out, world)
    in0, in1 = multi_gpu.image_planes_for_pieces(pieces[rank], size, size, off, out)
    slab = synth_volume_planes(size, in0, in1, 4321, dev)

    def step():
        return multi_gpu.detect_volume_rows_sharded(net, slab, in0, size, pieces, plans, NORM, DET["obj_min_dist"],
                                                    DET["smoothing_sigma"], (0, 0, 0), DET["buffer_sz"], DET["thd"])
    step()
    ms, _, out_d = timed(step, 1, sync)
    t = torch.tensor([ms], device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t)
    return {"workload": "unet_like2 (bf16) inference + exact-global voxel2obj on ONE synthetic %d^3 uint8 volume sharded over "
                        "%d GPUs, tile rows of the reference grid dealt evenly (BASELINE configs[2])" % (size, world),
            "value": size ** 3 / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "n_gpus": world,
            "detections": int(out_d["conf"].size), "tile_rows_per_rank": [sum(yb - ya for _, ya, yb, _ in m) for m in pieces],
            "planes_owned_per_rank": [p[1][1] - p[1][0] for p in plans]}


def leg_cfg5(dev, sync, world, rank, batch=64, steps=10):
    """BASELINE configs[4]: one data-parallel training step of vgg_like2 on synthetic minibatches (64 patches of 24^3 per
    GPU, already in HBM): forward / dgrad / wgrad on the tensor cores (bf16 hi/lo x3, fp32 accumulation), gradient
    all-reduce over NCCL at N > 1, Adam.  tools/bench_train.py is the stand-alone form with the stage split."""
    import torch
    from flypylib_b200 import fplmodels, fpltrain
    rf = 24
    model = fplmodels.vgg_like2(rf)[0]
    model.set_weights(seeded_weights("vgg_like2"))
    tr = fpltrain.Trainer(model, rf, batch, precision="tf32")
    g = torch.Generator(device=dev); g.manual_seed(77 + rank)
    x = torch.randn((batch, rf, rf, rf), generator=g, device=dev)
    y = (torch.rand(batch, generator=g, device=dev) < 0.5).to(torch.uint8)
    gb = batch * world
    it = [0]

    def step():
        it[0] += 1
        loss, _ = tr.forward_backward(x, y, gb, 1000 + it[0])
        tr.allreduce()
        tr.apply()
        return loss
    for _ in range(3):
        step()
    ms, _, loss = timed(step, steps, sync)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    d, fl = rf, 0.0
    for ci, (k, cin, cout) in enumerate(fplmodels._ARCH["vgg_like2"]["convs"]):
        d -= k - 1
        fl += 2.0 * k ** 3 * cin * cout * d ** 3
        if ci in (1, 3):
            d //= 2
    tr.close()
    return {"workload": "data-parallel training step, vgg_like2, %d patches of 24^3 per GPU, %d GPU(s) (BASELINE configs[4])"
                        % (batch, world), "ms_per_step": ms, "value": gb / (ms * 1e-3), "unit": "patches/s",
            "arithmetic": "tcgen05 bf16 hi/lo x3, fp32 accumulate", "n_gpus": world, "loss_sum_last_step": float(loss),
            "algorithmic_tflops": 3.0 * fl * gb / (ms * 1e-3) / 1e12,
            "round1_fp32_cuda_core_ms_per_step": 76.0}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from flypylib_b200 import fplobjdetect, multi_gpu, _lib

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

    net = build_net(args.model, args.precision, args.tile_mult)
    args.tile_mult = net.tile_multiplier
    ctx = _lib.context(local)
    size = args.size
    det = (DET["obj_min_dist"], DET["smoothing_sigma"], (0, 0, 0), DET["buffer_sz"], DET["thd"])

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    # ---- the workload: ONE synthetic size^3 volume.  N = 1: resident on the GPU.  N > 1: z-slab sharded over the
    # ranks (strong scaling): every rank holds the planes its slab needs (slab + 2*rf_offset halo)
    if world == 1:
        vol = synth_volume_device(size, 1234, dev)
        pred = torch.empty((size, size, size), dtype=torch.float32, device=dev)
        h2d_bytes = int(vol.numel())

        v2o_stats = {}

        def step():
            net.infer_device(vol, normalize=NORM, out=pred)
            out, st = fplobjdetect.voxel2obj_device(pred, *det, return_stats=True)
            v2o_stats.update(st)
            if st.get("two_tier_declined", 0) > 0 and "decline_info" not in v2o_stats:
                import ctypes
                info = (ctypes.c_longlong * 8)()
                _lib.lib().fpl_debug_v2o_decline_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
                code = _lib.lib().fpl_debug_v2o_decline_info(ctx.handle, info)
                v2o_stats["decline_info"] = [int(code)] + [int(v) for v in info]
            return out
    else:
        plans = multi_gpu.shard_plan(size, int(net.rf_offset[0]), net.slab_granularity(), world)
        (in0, in1), _own = plans[rank]
        vol = synth_volume_planes(size, in0, in1, 1234, dev)
        h2d_bytes = int(sum((p[0][1] - p[0][0]) * size * size for p in plans))

        def step():
            # forward: slab + receptive-field halo, no communication; detection: exact-global voxel2obj (halo planes
            # over NCCL P2P, all-reduced radix histograms, per-round all-gather of the selected points) -> the
            # single-GPU detection list on every rank
            return multi_gpu.detect_volume_sharded(net, vol, size, plans, NORM, *det)

    # the clock sampler (nvidia-smi -lms) is started BEFORE the warm-up: its start-up (NVML initialisation takes driver
    # locks for ~1 s) must not fall into the timed region; it keeps sampling through the timed steps
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        out = step()
    sync()
    l0 = ctx.launch_count()
    ctx.profile_begin()
    ms_per_step, _, out = timed(step, args.steps, sync)
    prof = ctx.profile_end()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    ms_per_step = max_over_ranks(ms_per_step)
    s2_stages = None
    if world > 1:       # one extra, untimed step with a device synchronisation at every stage boundary: where the
        #                 exact-global detection spends its time on rank 0 (the forward pass is the rest of the step)
        _o, st_ = multi_gpu.detect_volume_sharded(net, vol, size, plans, NORM, *det, return_stats='stages')
        s2_stages = {k: round(v, 3) for k, v in st_['stage_ms'].items()}       # 'setup' = the forward pass of this rank
        s2_stages['rounds'] = st_['rounds']
        per_rank = [None] * world                                               # load balance: forward ms / total of every rank
        dist.all_gather_object(per_rank, (round(st_['stage_ms'].get('setup', 0.0), 2), round(sum(st_['stage_ms'].values()), 2)))
        s2_stages['forward_and_total_ms_by_rank'] = per_rank
    n_det = int(out["conf"].size)
    import hashlib      # same list at every N (strong scaling on one volume, exact-global detection): compare across runs
    det_sha = hashlib.sha256(np.ascontiguousarray(out["locs"]).tobytes() + np.ascontiguousarray(out["conf"]).tobytes()).hexdigest()[:16]
    value = size ** 3 / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: pinned host uint8 volume -> H2D -> infer -> voxel2obj -> D2H detection list
    e2e = e2e_dropin = None
    if not args.no_e2e:
        host = torch.empty(tuple(vol.shape), dtype=torch.uint8).pin_memory()
        host.copy_(vol)
        if world == 1:
            dvol = torch.empty_like(vol)

            def e2e_step():
                # public API with a HOST volume: the H2D copy is pipelined behind the convolutions chunk by chunk
                net.infer_host(host, normalize=NORM, out=pred, image_dev=dvol)
                return fplobjdetect.voxel2obj_device(pred, *det)
        else:
            def e2e_step():
                vol.copy_(host, non_blocking=True)
                return step()
        e2e_step()                                          # untimed: sizes the buffers
        e_ms, e_wall, out_e = timed(e2e_step, args.steps, sync)
        e_ms = max_over_ranks(max(e_ms, e_wall))
        e2e = {"value": size ** 3 / (e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": int(out_e["conf"].size * 32)}
        del host
        if world == 1 and not args.no_extras and args.precision == "bf16":
            # the literal reference call sequence (fplnetwork.py:136, fplobjdetect.py:132): normalised float32 ndarray
            # -> infer -> float32 ndarray -> voxel2obj(ndarray) -> dict; pageable host memory, 4 B/voxel each way
            try:
                del dvol
                img = ((vol.cpu().numpy().astype(np.float32) - np.float32(NORM[0])) / np.float32(NORM[1]))

                def dropin_step():
                    p = net.infer(img)
                    return fplobjdetect.voxel2obj(p, *det)
                dropin_step()
                t0 = time.perf_counter()
                out_d = dropin_step()
                d_ms = (time.perf_counter() - t0) * 1e3
                e2e_dropin = {"value": size ** 3 / (d_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": d_ms,
                              "h2d_bytes_per_step": int(8 * size ** 3), "d2h_bytes_per_step": int(4 * size ** 3 + out_d["conf"].size * 32),
                              "call": "voxel2obj(FplNetwork.infer(float32 ndarray)) with pageable numpy arrays both ways",
                              "identical_to_device_path": bool(np.array_equal(out_d["conf"], out["conf"]) and
                                                               np.array_equal(out_d["locs"], out["locs"]))}
                del img
            except Exception as e:      # noqa: BLE001
                e2e_dropin = {"error": repr(e)[:200]}

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
                "kernel": "conv_fused12_kernel + conv_umma_kernel<3> (tcgen05 implicit GEMM, 3x3x3 convolutions)"
                          + (" on rank 0" if world > 1 else ""),
                "launches": int(cnt3 / max(1, args.steps)), "ms_per_step": ms3 / args.steps,
                "peak_source": "%s bf16 sustained" % pk["src"]}
    families = {k: {"ms_per_step": v[0] / args.steps, "launch_groups": int(v[2] / max(1, args.steps))}
                for k, v in prof.items() if v[2]}
    g_ms, g_work, _ = prof["gauss"]
    if g_ms > 0:
        families["gauss"]["hbm_frac_algorithmic"] = (g_work / (g_ms * 1e-3) / 1e9) / pk["hbm"]
    if world == 1:
        how = ("synthetic %d^3 uint8 EM volume resident in HBM, random-init weights, reference tile grid evaluated as "
               "z-slab tiles of %d reference layers" % (size, args.tile_mult))
    else:
        how = ("ONE synthetic %d^3 uint8 EM volume z-slab sharded over %d GPUs (cuts every %d planes, 2*rf_offset halo, no "
               "forward communication), exact-global voxel2obj: same detection list as N=1 on every rank"
               % (size, world, net.slab_granularity()))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "%s (%s, tcgen05 implicit GEMM) inference + voxel2obj(r=27, sigma=5, buffer=15) on a %s"
                                   % (args.model, args.precision, how),
                       "l2": "inputs larger than L2 (%.2f GiB uint8 volume, %.1f GiB probability map per GPU and step)"
                             % (vol.numel() / 2 ** 30, 4.0 * size ** 3 / world / 2 ** 30),
                       "detections_per_step": n_det, "detections_sha256_16": det_sha,
                       "voxel2obj_path": (v2o_stats.get("path"), v2o_stats.get("two_tier_declined"), v2o_stats.get("decline_info")) if world == 1 else "exact-global (slab sessions)"},
            "roofline": roofline, "families": families, "gpu_launches": int(launches / max(1, args.steps)),
            "clocks": clocks, "e2e": e2e}
    if e2e_dropin is not None:
        line["e2e_dropin"] = e2e_dropin
    if s2_stages is not None:
        line["detection_stages_ms_rank0"] = s2_stages

    # ---- extra legs (never allowed to take the main line down): the other BASELINE configs, driver-visible
    def emit():
        if rank == 0:
            print(json.dumps(line), flush=True)

    if not args.no_extras and args.model == "vgg_like2" and args.precision == "bf16" and size == 1024:
        import threading
        watchdog = threading.Timer(240.0, lambda: (line.setdefault("extras_error", "extra legs timed out"), emit(),
                                                   os._exit(0)))
        watchdog.daemon = True
        watchdog.start()
        del vol, out
        if world == 1:
            del pred
        net.infer_network.close(); net.train_single.close()
        torch.cuda.empty_cache()
        ctx.release_workspace()
        try:
            if world == 1:
                line["cfg1"] = leg_cfg1(dev, sync)
                ctx.release_workspace(); torch.cuda.empty_cache()
                line["cfg4"] = leg_cfg4(dev, sync, ctx, pk)
                ctx.release_workspace(); torch.cuda.empty_cache()
                line["cfg5"] = leg_cfg5(dev, sync, world, rank)
            else:
                line["cfg3"] = leg_cfg3(dev, sync, world, rank)
                line["cfg5"] = leg_cfg5(dev, sync, world, rank)
        except Exception as e:      # noqa: BLE001
            line["extras_error"] = repr(e)[:300]
        watchdog.cancel()
        if world > 1 and "extras_error" in line:        # ranks may be out of step: no further collectives
            emit()
            os._exit(0)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.model)
    emit()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
