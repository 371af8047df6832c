"""GPU parity: CUDA voxel2obj (through the C ABI) against the oracle and the reference goldens.
Bit-exact: smoothed map, threshold, detection coordinates, confidences and their order."""
import ctypes
import os

import numpy as np
import pytest

from oracle import voxel2obj_oracle as O
from tests.golden import cases

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel2obj_golden.npz"))


def _stages(pred, r, sigma, off, buf, thd, force_generic=False):
    """smooth / threshold / detect through the staged ABI calls."""
    import torch
    from flypylib_b200 import _lib, fplobjdetect as P
    dev = torch.device("cuda", 0)
    ctx = _lib.context(0)
    lib = _lib.lib()
    lib.fpl_debug_force_generic_gauss.argtypes = [ctypes.c_int]
    lib.fpl_debug_force_generic_gauss(1 if force_generic else 0)
    try:
        d_pred = torch.from_numpy(pred).to(dev)
        d_s = torch.empty_like(d_pred)
        p, keep = P._make_params(pred.shape, r, sigma, off, buf, thd)
        Z, Y, X = pred.shape
        st = _lib.current_stream_ptr(0)
        _lib.check(lib.fpl_v2o_smooth(ctx.handle, d_pred.data_ptr(), Z, Y, X, ctypes.byref(p),
                                      d_s.data_ptr(), st), "smooth")
        hout = (ctypes.c_double * 4)()
        _lib.check(lib.fpl_v2o_threshold(ctx.handle, d_s.data_ptr(), Z, Y, X, ctypes.byref(p), hout, st),
                   "threshold")
        cap = P._default_capacity(pred.shape, r)
        dets = torch.empty((cap, 4), dtype=torch.float64, device=dev)
        cnt = ctypes.c_int64()
        stats = (ctypes.c_int64 * 8)()
        _lib.check(lib.fpl_v2o_detect(ctx.handle, d_s.data_ptr(), Z, Y, X, ctypes.byref(p), hout[0],
                                      dets.data_ptr(), cap, ctypes.byref(cnt), stats, st), "detect")
        rows = dets[:cnt.value].cpu().numpy()
        return d_s.cpu().numpy(), hout[0], rows, list(stats)
    finally:
        lib.fpl_debug_force_generic_gauss(0)


@pytest.mark.parametrize("case", cases.VOXEL2OBJ_CASES, ids=[c[0] for c in cases.VOXEL2OBJ_CASES])
@pytest.mark.parametrize("generic", [False, True], ids=["fast", "generic"])
def test_stages_bit_exact_vs_oracle_and_golden(case, generic):
    name, shape, seed, kind, r, sigma, thd, buf, off = case
    pred = cases.prob_map(shape, seed, kind)
    s_gpu, t_gpu, rows, stats = _stages(pred, r, sigma, off, buf, thd, force_generic=generic)
    want, s_ref, t_ref = O.voxel2obj(pred, r, sigma, off, buf, thd, impl="c", return_intermediates=True)
    interior = s_ref[r:r + shape[0], r:r + shape[1], r:r + shape[2]] if r > 0 else s_ref
    assert np.array_equal(s_gpu.view(np.uint32), np.ascontiguousarray(interior).view(np.uint32)), \
        "smoothed map not bit-exact"
    assert t_gpu == float(t_ref)
    assert np.array_equal(rows[:, :3], want["locs"])
    assert np.array_equal(rows[:, 3], want["conf"])
    # and against the reference's own output
    assert np.array_equal(rows[:, :3], GOLD[name + "/locs"])
    assert np.array_equal(rows[:, 3], GOLD[name + "/conf"])


@pytest.fixture(autouse=True)
def _no_two_tier_backoff():
    """fpl_voxel2obj backs off from the two-tier path after data-dependent declines (adaptive, per context); tests
    must not inherit that state from the maps of the previous test."""
    from flypylib_b200 import _lib
    lib = _lib.lib()
    lib.fpl_debug_v2o_decline_reason.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.fpl_debug_v2o_decline_reason(_lib.context(0).handle, 1)
    yield


@pytest.fixture(params=["default", "fused", "classic"])
def v2o_path(request):
    """fpl_voxel2obj has three detection paths: the two-tier one (fp32 smoothing with a proven bound, exact values
    only where a decision needs them; the default whenever the map qualifies), the fused exact one (exact smoothing,
    three dense passes, NMS on a candidate superset) and the classic exact one it falls back to when the superset
    does not fit its lists.  Every drop-in parity test runs through all three."""
    from flypylib_b200 import _lib
    lib = _lib.lib()
    lib.fpl_debug_v2o_classic.argtypes = [ctypes.c_int]
    lib.fpl_debug_v2o_classic({"default": 0, "classic": 1, "fused": 2}[request.param])
    yield request.param
    lib.fpl_debug_v2o_classic(0)


@pytest.mark.parametrize("case", cases.VOXEL2OBJ_CASES, ids=[c[0] for c in cases.VOXEL2OBJ_CASES])
def test_dropin_voxel2obj_matches_golden(case, v2o_path):
    from flypylib_b200 import fplobjdetect
    name, shape, seed, kind, r, sigma, thd, buf, off = case
    pred = cases.prob_map(shape, seed, kind)
    out = fplobjdetect.voxel2obj(pred, r, sigma, off, buf, thd)
    assert out["locs"].dtype == np.float64 and out["conf"].dtype == np.float64
    assert out["locs"].shape == GOLD[name + "/locs"].shape
    assert np.array_equal(out["locs"], GOLD[name + "/locs"])
    assert np.array_equal(out["conf"], GOLD[name + "/conf"])


@pytest.mark.parametrize("shape,kind,seed", [((200, 190, 210), "blobs", 21), ((192, 192, 192), "uniform", 22),
                                             ((256, 256, 256), "blobs", 23), ((180, 200, 170), "ties", 24)])
def test_medium_volumes_vs_c_oracle(shape, kind, seed, v2o_path):
    """Reference parameters (r=27, sigma=5) at sizes the C oracle finishes in seconds."""
    from flypylib_b200 import fplobjdetect
    pred = cases.prob_map(shape, seed, kind)
    got, st = fplobjdetect.voxel2obj_device(__import__("torch").from_numpy(pred).cuda(), 27, 5, (0, 0, 0), 30, 0,
                                            return_stats=True)
    want, s, t = O.voxel2obj(pred, 27, 5, (0, 0, 0), 30, 0, impl="c", return_intermediates=True)
    assert st["threshold"] == float(t)
    assert np.array_equal(got["locs"], want["locs"])
    assert np.array_equal(got["conf"], want["conf"])


def test_nan_and_saturation_edge_cases(v2o_path):
    from flypylib_b200 import fplobjdetect
    a = cases.prob_map((40, 40, 40), 3, "blobs")
    a[5, 6, 7] = np.nan
    out = fplobjdetect.voxel2obj(a, 4, 1.0)
    assert out["locs"].shape == (0, 3) and out["conf"].shape == (0,)
    out = fplobjdetect.voxel2obj(cases.prob_map((32, 32, 32), 8, "saturated"), 4, 1.0)
    assert out["locs"].shape == (0, 3)


def test_large_volume_properties():
    """512^3 with reference parameters: size-independent properties (no oracle at this size):
    pairwise distance > r, confidences sorted, every detection above the threshold, idempotent."""
    import torch
    from flypylib_b200 import fplobjdetect
    from flypylib_b200 import _lib
    pred = cases.prob_map((512, 512, 512), 77, "blobs")
    d = torch.from_numpy(pred).cuda()
    out, st = fplobjdetect.voxel2obj_device(d, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
    out2 = fplobjdetect.voxel2obj_device(d, 27, 5, (0, 0, 0), 30, 0)
    assert np.array_equal(out["locs"], out2["locs"]) and np.array_equal(out["conf"], out2["conf"])
    # fused path == classic path (exact threshold first, then candidates = smooth > threshold)
    lib = _lib.lib()
    lib.fpl_debug_v2o_classic.argtypes = [ctypes.c_int]
    lib.fpl_debug_v2o_classic(1)
    try:
        out3, st3 = fplobjdetect.voxel2obj_device(d, 27, 5, (0, 0, 0), 30, 0, return_stats=True)
    finally:
        lib.fpl_debug_v2o_classic(0)
    assert st3["threshold"] == st["threshold"]
    assert np.array_equal(out["locs"], out3["locs"]) and np.array_equal(out["conf"], out3["conf"])
    conf = out["conf"]
    assert conf.size > 100
    assert np.all(np.diff(conf) <= 0)
    assert np.all(conf > st["threshold"])
    from scipy.spatial import cKDTree
    pairs = cKDTree(out["locs"]).query_pairs(27.0)
    assert len(pairs) == 0
    lo = out["locs"].min(0); hi = out["locs"].max(0)
    assert np.all(lo >= 30) and np.all(hi < 512 - 30)


def test_slab_sharded_detection_matches_reference_substack_semantics():
    """N>1 detection path (emulated ranks on one GPU): per-slab voxel2obj with a buffer, detections in the
    buffer dropped, lists merged == the same procedure carried out with the oracle (reference function
    restated) on every slab.  This is full_roi_inference's substack semantics (fplobjdetect.py:841-986)."""
    import torch
    from flypylib_b200 import multi_gpu
    pm = cases.prob_map((200, 120, 110), 31, "uniform")
    d = torch.from_numpy(pm).cuda()
    world, r, sigma, buf = 3, 12, 2.0, 16
    parts, want_parts = [], []
    for rank in range(world):
        parts.append(multi_gpu.slab_detections(lambda lo, hi: d[lo:hi].contiguous(), 200, rank, world, r, sigma, buf))
        z0, z1 = multi_gpu.partition_layers(200, world)[rank]
        lo, hi = max(0, z0 - buf), min(200, z1 + buf)
        o = O.voxel2obj(pm[lo:hi], r, sigma, (0, 0, lo), 0, 0, impl="c")
        rows = np.concatenate([o["locs"], o["conf"][:, None]], 1)
        want_parts.append(rows[(rows[:, 2] >= z0) & (rows[:, 2] < z1)])
    got = multi_gpu.merge_detections(parts)
    want = multi_gpu.merge_detections(want_parts)
    assert got["conf"].size > 20
    assert np.array_equal(got["locs"], want["locs"]) and np.array_equal(got["conf"], want["conf"])


def test_randomised_parameters_vs_oracle(v2o_path):
    """Seeded sweep over shapes / radii / sigmas / thd / buffers / offsets / map kinds (the reference has no
    tests of its own): CUDA result must equal the C oracle bit-for-bit every time."""
    from flypylib_b200 import fplobjdetect
    rng = np.random.default_rng(2024)
    kinds = ["blobs", "uniform", "ties"]
    for it in range(24):
        shape = tuple(int(v) for v in rng.integers(12, 72, 3))
        r = int(rng.integers(1, 13))
        sigma = float(rng.choice([0.0, 0.8, 1.0, 1.5, 2.0, 2.5, 3.3, 4.0, 5.0]))
        thd = float(rng.choice([0, 0, 0.02, 0.3]))
        buf = int(rng.integers(0, 4)) if it % 2 else tuple(int(v) for v in rng.integers(0, 4, 3))
        off = tuple(int(v) for v in rng.integers(-50, 50, 3))
        pm = cases.prob_map(shape, 1000 + it, kinds[it % 3], peaks_per_50cube=40.0)
        got = fplobjdetect.voxel2obj(pm, r, sigma, off, buf, thd)
        want = O.voxel2obj(pm, r, sigma, off, buf, thd, impl="c")
        ctx = "case %d shape %s r %d sigma %g thd %g buf %s off %s" % (it, shape, r, sigma, thd, buf, off)
        assert np.array_equal(got["locs"], want["locs"]), ctx
        assert np.array_equal(got["conf"], want["conf"]), ctx


@pytest.mark.parametrize("margin", [0, 64, 4096, 1 << 28], ids=["exact-only", "default", "wide", "all-recomputed"])
def test_gaussian_certified_chain_is_bit_exact(margin):
    """The Gaussian passes evaluate a fused (FMA) chain first and prove per output that it rounds to the same
    float32 as SciPy's separately rounded chain; outputs that fail the certificate are recomputed exactly.
    Whatever the margin (0 = fast chain off, 2^28 = every output recomputed), the smoothed map must be
    bit-identical to the oracle -- including negative inputs (certificate not applicable) and huge/tiny values."""
    from flypylib_b200 import _lib
    lib = _lib.lib()
    lib.fpl_debug_gauss_cert.argtypes = [ctypes.c_int]
    rng = np.random.default_rng(5)
    maps = [cases.prob_map((70, 66, 90), 11, "blobs"), cases.prob_map((64, 64, 64), 12, "ties"),
            (rng.standard_normal((48, 50, 52)) * 3).astype(np.float32),                  # negative inputs
            (rng.random((40, 40, 70), dtype=np.float32) * np.float32(1e-38)),             # float32 subnormal results
            (rng.random((40, 40, 70), dtype=np.float32) * np.float32(3e38))]              # near overflow
    lib.fpl_debug_gauss_cert(margin)
    try:
        for i, pm in enumerate(maps):
            for r, sigma in ((6, 5.0), (5, 1.5)):
                s_gpu, _, _, _ = _stages(pm, r, sigma, (0, 0, 0), 0, 0)
                _, s_ref, _ = O.voxel2obj(pm, r, sigma, (0, 0, 0), 0, 0, impl="c", return_intermediates=True)
                interior = np.ascontiguousarray(s_ref[r:r + pm.shape[0], r:r + pm.shape[1], r:r + pm.shape[2]])
                assert np.array_equal(s_gpu.view(np.uint32), interior.view(np.uint32)), (i, r, sigma)
    finally:
        lib.fpl_debug_gauss_cert(0)


@pytest.mark.parametrize("shape,kind,mode", [((416, 512, 448), "blobs", 0), ((416, 512, 448), "blobs", 1),
                                             ((416, 512, 448), "blobs", 2), ((352, 384, 400), "uniform", 0)])
def test_scale_runner_vs_c_oracle(shape, kind, mode):
    """tools/check_v2o_scale.py (the runner behind profiles/r02_v2o_scale_*.json: the 1024^3 bench map and a map with
    more than 2^32 voxels) on shapes the C oracle finishes in well under a minute: reference parameters, threshold,
    list, order and confidences bit for bit, on the default and on the classic path."""
    from tools import check_v2o_scale
    res = check_v2o_scale.compare(shape, kind, 7, 27, 5.0, 15, 0.0, mode=mode)
    assert res["detections_oracle"] > 20
    if mode == 0 and kind == "uniform":     # i.i.d. noise smoothed with sigma 5 crowds within 1 % of its mean: the two-tier
        assert res["gpu_path"] in ("two-tier", "fused-exact")     # path may hand it to the exact path (cost guard)
    else:
        assert res["gpu_path"] == {0: "two-tier", 1: "classic-exact", 2: "fused-exact"}[mode], res["gpu_path"]
    assert res["threshold_identical"] and res["locs_identical"] and res["conf_identical"], res


@pytest.mark.parametrize("shape,kind,r,sigma,thd", [
    ((128, 128, 128), "blobs", 27, 5.0, 0), ((96, 100, 132), "uniform", 27, 5.0, 0), ((120, 90, 101), "blobs", 9, 2.0, 0),
    ((100, 96, 96), "uniform", 8, 4.0, 0.51), ((90, 110, 70), "blobs", 6, 1.5, 0.02), ((64, 64, 200), "uniform", 5, 1.0, 0),
])
def test_two_tier_path_runs_and_is_bit_exact(shape, kind, r, sigma, thd):
    """The default path on maps that qualify (non-negative, lw <= r) IS the two-tier path (stats say so), and its
    threshold / list / order / confidences equal the C oracle's bit for bit; odd x extents take the scalar staging."""
    import torch
    from flypylib_b200 import fplobjdetect
    pm = cases.prob_map(shape, 5, kind, peaks_per_50cube=20.0)
    got, st = fplobjdetect.voxel2obj_device(torch.from_numpy(pm).cuda(), r, sigma, (0, 0, 0), 3, thd, return_stats=True)
    assert st["path"] == "two-tier", st
    want, s, t = O.voxel2obj(pm, r, sigma, (0, 0, 0), 3, thd, impl="c", return_intermediates=True)
    assert st["threshold"] == float(t)
    assert thd or want["conf"].size > 3
    assert np.array_equal(got["locs"], want["locs"]) and np.array_equal(got["conf"], want["conf"])


def test_two_tier_path_declines_what_it_cannot_certify():
    """Negative or non-finite inputs, lw > r, ties / plateaus: the exact paths take over (and stay bit-exact)."""
    import torch
    from flypylib_b200 import fplobjdetect
    rng = np.random.default_rng(3)
    neg = (rng.standard_normal((64, 64, 64)) * 0.3).astype(np.float32)
    for pm, r, sigma in [(neg, 6, 1.5), (cases.prob_map((64, 64, 64), 4, "blobs"), 3, 4.0),
                         (cases.prob_map((80, 80, 80), 4, "ties"), 6, 1.5), (cases.prob_map((64, 64, 64), 8, "saturated"), 4, 1.0)]:
        got, st = fplobjdetect.voxel2obj_device(torch.from_numpy(pm).cuda(), r, sigma, (0, 0, 0), 0, 0, return_stats=True)
        want = O.voxel2obj(pm, r, sigma, (0, 0, 0), 0, 0, impl="c")
        assert np.array_equal(got["locs"], want["locs"]) and np.array_equal(got["conf"], want["conf"]), st
    got, st = fplobjdetect.voxel2obj_device(torch.from_numpy(neg).cuda(), 6, 1.5, (0, 0, 0), 0, 0, return_stats=True)
    assert st["path"] != "two-tier"
