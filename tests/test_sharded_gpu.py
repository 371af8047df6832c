"""GPU: the north-star multi-GPU split and the benchmarked tile schedule.

ONE volume, z-slab sharded at the network's grid granularity (``multi_gpu.shard_plan``), every rank evaluating its
slab + 2*rf_offset halo through ``fpl_net_infer_slab``; then exact-global ``voxel2obj``.  The ranks are emulated in
one process on one GPU (the forward pass needs no communication; ``_LocalCollectives`` turns the detection
collectives into tensor operations).  Everything must equal the single-GPU result bit for bit.  Replaces
flypylib/multi_gpu.py:20-61 + fplnetwork.py:130-134,175-176.
"""
import numpy as np
import pytest

from oracle import models_oracle as M
from tests.golden import cases

pytestmark = pytest.mark.gpu
NORM = (128.0, 33.0)


def _net(arch, precision, seed, tile_mult):
    from flypylib_b200 import fplmodels, fplnetwork
    net = fplnetwork.FplNetwork(getattr(fplmodels, arch))
    net.train_single.set_weights(M.random_weights(arch, seed=seed))
    net.set_precision(precision)
    net._set_infer()
    net.tile_multiplier = tile_mult
    return net


@pytest.mark.parametrize("arch,shape,world,tile_mult", [
    ("vgg_like2", (300, 200, 185), 3, 4),        # cuts every 4 planes: 70 groups -> 24/23/23
    ("vgg_like2", (120, 130, 101), 5, 1),        # slabs thinner than a reference tile
    ("vgg_like", (150, 120, 110), 2, 4),
    ("unet_like2", (290, 100, 120), 3, 1),       # 4 tile layers of 82 -> 2/1/1, cuts on the tile grid only
    ("unet_like2", (100, 105, 100), 2, 1),       # one layer: the second rank has no work
])
def test_sharded_infer_equals_whole_volume(arch, shape, world, tile_mult):
    import torch
    from flypylib_b200 import multi_gpu
    net = _net(arch, "bf16", 21, tile_mult)
    u8 = torch.from_numpy(cases.em_volume(shape, seed=6)).cuda()
    want = net.infer_device(u8, normalize=NORM)
    Z = shape[0]
    plans = multi_gpu.shard_plan(Z, int(net.rf_offset[0]), net.slab_granularity(), world)
    got = torch.full(shape, float("nan"), dtype=torch.float32, device="cuda")
    for (in0, in1), (own0, own1) in plans:
        if in1 <= in0:
            continue
        # (a) into a fresh per-rank buffer: exactly the owned planes come back
        own, first, last = net.infer_slab_device(u8[in0:in1], Z, in0, normalize=NORM)
        assert (first, last) == (own0, own1) and own.shape[0] == own1 - own0
        assert torch.equal(own, want[own0:own1])
        # (b) into the shared prediction volume: only the owned planes are touched
        net.infer_slab_device(u8[in0:in1], Z, in0, normalize=NORM, pred=got, pred_z0=0)
    assert not torch.isnan(got).any()
    assert torch.equal(got, want)


def test_slab_api_rejects_cuts_off_the_grid():
    import torch
    from flypylib_b200 import _lib
    net = _net("vgg_like2", "bf16", 2, 4)
    u8 = torch.zeros((64, 64, 64), dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.FplError, match="multiple of 4"):
        net.infer_slab_device(u8[:50], 200, 6, normalize=NORM)
    with pytest.raises(_lib.FplError, match="inner slab"):
        net.infer_slab_device(u8[:50], 200, 8, normalize=NORM)       # 50 - 20 is not a multiple of 4
    unet = _net("unet_like2", "bf16", 2, 1)
    with pytest.raises(_lib.FplError, match="multiple of 82"):
        unet.infer_slab_device(u8, 300, 80, normalize=NORM)


@pytest.mark.parametrize("arch,shape,world,r,sigma,buf", [
    ("vgg_like2", (260, 150, 140), 3, 9, 2.0, 6),
    ("unet_like2", (270, 100, 100), 2, 8, 2.0, 5),
])
def test_sharded_path_equals_single_gpu_detections(arch, shape, world, r, sigma, buf):
    """infer (sharded) + voxel2obj_global == infer (whole) + voxel2obj: same list, same order, same bits."""
    import torch
    from flypylib_b200 import multi_gpu, fplobjdetect
    net = _net(arch, "bf16", 5, 4 if arch.startswith("vgg") else 1)
    u8 = torch.from_numpy(cases.em_volume(shape, seed=12)).cuda()
    want = fplobjdetect.voxel2obj_device(net.infer_device(u8, normalize=NORM), r, sigma, (1, 2, 3), buf, 0)
    Z = shape[0]
    plans = multi_gpu.shard_plan(Z, int(net.rf_offset[0]), net.slab_granularity(), world)
    slabs, ranges = [], []
    for (in0, in1), (own0, own1) in plans:
        if in1 > in0:
            own, _, _ = net.infer_slab_device(u8[in0:in1], Z, in0, normalize=NORM)
        else:
            own = torch.zeros((0,) + shape[1:], dtype=torch.float32, device="cuda")
        slabs.append(own)
        ranges.append((own0, own1))
    got = multi_gpu.voxel2obj_global(slabs, ranges, Z, r, sigma, (1, 2, 3), buf, 0,
                                     coll=multi_gpu._LocalCollectives(world))
    assert want["conf"].size > 5
    assert np.array_equal(got["locs"], want["locs"]) and np.array_equal(got["conf"], want["conf"])


def test_infer_host_chunks_write_in_place():
    """infer_host (pinned host volume, H2D pipelined chunk by chunk, every chunk a slab written straight into the
    result) == infer_device on the resident volume."""
    import torch
    net = _net("vgg_like2", "bf16", 9, 4)
    host = torch.from_numpy(cases.em_volume((430, 120, 110), seed=3)).pin_memory()      # 6 reference layers
    want = net.infer_device(host.cuda(), normalize=NORM)
    got = net.infer_host(host, normalize=NORM, chunk_layers=2)
    assert torch.equal(got, want)
    got4 = net.infer_host(host, normalize=NORM, chunk_layers=4)
    assert torch.equal(got4, want)


# ---------------------------------------------------------------------------------------------------
# the schedule bench.py times: tile_multiplier = 4 z-slab tiles (full x/y extent)
# ---------------------------------------------------------------------------------------------------
def test_bench_tile_schedule_equals_reference_grid():
    """vgg_like2 bf16, tile_multiplier=4 on a volume of 5 reference layers whose edges are not multiples of the
    tile pitch == the reference tile grid (tile_multiplier=1), bit for bit."""
    import torch
    net = _net("vgg_like2", "bf16", 5, 1)
    u8 = torch.from_numpy(cases.em_volume((420, 330, 350), seed=9)).cuda()
    ref = net.infer_device(u8, normalize=NORM)
    net.tile_multiplier = 4
    got = net.infer_device(u8, normalize=NORM)
    assert torch.equal(ref, got)
    assert float(got[10:-10, 10:-10, 10:-10].min()) > 0.0 and not got[:10].any()
    # tail slab trimmed to the planes that are left (rounded up to rf_stride): 431 planes -> 320 + 92 (91 needed)
    for zcut in (431, 346, 118):
        sub = u8[:zcut].contiguous()
        net.tile_multiplier = 1
        ref = net.infer_device(sub, normalize=NORM)
        net.tile_multiplier = 4
        assert torch.equal(ref, net.infer_device(sub, normalize=NORM)), zcut


def test_bf16_config_tile_and_slab_tile_vs_float64():
    """bf16 tcgen05 path against the float64 restatement at the sizes that are benchmarked: one 100^3
    configuration tile (vgg_like2 infer_sz) and one z-slab-shaped tile (44 x 120 x 120 volume evaluated as a
    single slab tile vs the oracle's reference tiling)."""
    import torch
    arch = "vgg_like2"
    w = M.random_weights(arch, seed=11)
    from flypylib_b200 import fplmodels
    model, _, _, _ = fplmodels.vgg_like2(100)
    model.upsample_output = True
    model.set_precision("bf16")
    model.set_weights(w)
    x = np.random.default_rng(100).standard_normal((1, 100, 100, 100)).astype(np.float32)
    got = model.predict(x[..., None], batch_size=1)[..., 0]
    want = M.forward(arch, w, x)
    assert got.shape == want.shape == (1, 80, 80, 80)
    e = np.abs(got.astype(np.float64) - want).max()
    assert e < 4e-3, "bf16 100^3 tile vs float64 oracle: %g" % e
    # slab-shaped tile through the volume API
    net = _net(arch, "bf16", 11, 4)
    img = ((cases.em_volume((44, 120, 120), seed=4).astype(np.float32) - 128.0) / 33.0).astype(np.float32)
    got_v = net.infer_device(torch.from_numpy(img).cuda()).cpu().numpy()
    want_v = M.infer_tiler(img, M.TorchNet(arch, w, dtype=torch.float64), net.infer_sz, net.rf_offset, n_gpu=1)
    e = np.abs(got_v.astype(np.float64) - want_v.astype(np.float64)).max()
    assert e < 4e-3, "bf16 slab tile vs float64 oracle tiling: %g" % e


def test_row_sharded_unet_equals_whole_volume():
    """U-Net with the (z, y) tile rows dealt to the ranks (multi_gpu.row_plan): every piece evaluated as an independent
    block on the reference tile grid, assembled == whole-volume infer, bit for bit."""
    import torch
    from flypylib_b200 import multi_gpu
    net = _net("unet_like2", "bf16", 7, 1)
    shape = (270, 200, 190)                                  # 4 layers x 3 rows of tiles
    u8 = torch.from_numpy(cases.em_volume(shape, seed=8)).cuda()
    want = net.infer_device(u8, normalize=NORM)
    Z, Y, X = shape
    off, out = 9, 82
    pieces, plans = multi_gpu.row_plan(Z, Y, off, out, 5)
    got = torch.zeros(shape, dtype=torch.float32, device="cuda")
    for mine in pieces:
        for piece in mine:
            (zr, yr), ((pz0, pz1), (py0, py1)) = multi_gpu.piece_geometry(piece, Z, Y, off, out)
            sub = net.infer_device(u8[zr[0]:zr[1], yr[0]:yr[1]].contiguous(), normalize=NORM)
            got[pz0:pz1, py0:py1] = sub[off:off + pz1 - pz0, off:off + py1 - py0]
    assert torch.equal(got, want)
