"""CPU: the oracle (oracle/voxel2obj_oracle.py, oracle/voxel2obj_c.c) against the golden vectors
produced by the unmodified reference (tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import voxel2obj_oracle as O
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel2obj_golden.npz"))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("case", cases.VOXEL2OBJ_CASES, ids=[c[0] for c in cases.VOXEL2OBJ_CASES])
@pytest.mark.parametrize("impl", ["numpy", "numpy-literal", "c"])
def test_oracle_matches_reference_golden(case, impl):
    name, shape, seed, kind, r, sigma, thd, buf, off = case
    if impl == "numpy-literal" and np.prod(shape) > 64 ** 3:
        pytest.skip("literal O(K*C) loop only on small cases")
    if impl == "numpy" and np.prod(shape) > 130 ** 3:
        pytest.skip("numpy path on small/medium cases")
    pred = cases.prob_map(shape, seed, kind)
    out, s, t = O.voxel2obj(pred, r, sigma, off, buf, thd, impl=impl, return_intermediates=True)
    assert np.array_equal(_sha(s), GOLD[name + "/smooth_sha256"]), "smoothed map differs"
    assert np.asarray(t).dtype == GOLD[name + "/thresh"].dtype and t == GOLD[name + "/thresh"]
    assert out["locs"].dtype == np.float64 and out["conf"].dtype == np.float64
    assert out["locs"].shape == GOLD[name + "/locs"].shape
    assert np.array_equal(out["locs"], GOLD[name + "/locs"])
    assert np.array_equal(out["conf"], GOLD[name + "/conf"])


def test_percentile_restatement_matches_numpy():
    rng = np.random.default_rng(5)
    for n in [1, 2, 3, 10, 33, 100, 101, 1000, 4097, 65537, 300001]:
        for _ in range(3):
            a = rng.standard_normal(n).astype(np.float32)
            if n > 10:
                a[rng.integers(0, n, n // 3)] = 0  # ties / zeros
            got = O.percentile_f32(a, 97)
            want = np.percentile(a, 97)
            assert got.dtype == want.dtype == np.float32
            assert got == want, (n, got, want)


def test_percentile_plan_large_n_is_float32_index():
    # 2048^3 padded with r=27: gamma == 0 and the rank is the float32-rounded index
    n = 2102 ** 3
    lo, hi, g = O.percentile_plan(n)
    assert g == 0 and lo == hi == 9008861184


def test_product_host_scalars_match_oracle():
    """The host shim derives taps / ranks / thd promotion itself; they must equal the oracle's."""
    from flypylib_b200 import fplobjdetect as P
    for sigma in [0.0, 1.0, 1.5, 2.0, 4.0, 5.0]:
        w, lw = P._gaussian_taps(sigma)
        if sigma == 0:
            assert w is None and lw == -1
            continue
        w2, lw2 = O.gaussian_weights(sigma)
        assert lw == lw2 and np.array_equal(w, w2)
    for n in [1, 5, 1000, 29791000, 150 ** 3, 310 ** 3, 1078 ** 3, 2102 ** 3]:
        assert P._percentile_plan(n) == tuple(
            float(v) if i == 2 else v for i, v in enumerate(O.percentile_plan(n)))
    assert P._promote_thd(0) == 0.0
    assert P._promote_thd(0.05) == float(np.float32(0.05))
    assert P._promote_thd(np.float64(0.05)) == 0.05


def test_seg_branch_of_the_oracle_matches_the_unmodified_reference():
    """oracle.voxel2obj_seg (seg / seg_dilate / seg_sz_thd / seg_force, fplobjdetect.py:161-224) == goldens produced by
    the unmodified reference (tests/golden/make_golden.py seg)."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "voxel2obj_seg_golden.npz"))
    for name, shape, seed, kind, r, sigma, thd, buf, off, dil, szt, force in cases.VOXEL2OBJ_SEG_CASES[:5]:
        pred = cases.prob_map(shape, seed, kind)
        seg = cases.segmentation(shape, seed + 100)
        out = O.voxel2obj_seg(pred, r, sigma, off, buf, thd, seg=seg, seg_dilate=dil, seg_sz_thd=szt, seg_force=force)
        assert np.array_equal(out["locs"], gold[name + "/locs"]), name
        assert np.array_equal(out["conf"], gold[name + "/conf"]), name
