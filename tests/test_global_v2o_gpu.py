"""GPU: exact multi-GPU voxel2obj (multi_gpu.voxel2obj_global, semantics S2 of SURVEY 8e).  All ranks are run in
one process on one GPU (`_LocalCollectives`: the collectives become tensor operations); the result must be
bit-identical to ONE voxel2obj call on the whole map, for any number and thickness of slabs."""
import numpy as np
import pytest

from tests.golden import cases

pytestmark = pytest.mark.gpu


def _run(pm, world, r, sigma, off, buf, thd, cuts=None):
    import torch
    from flypylib_b200 import multi_gpu, fplobjdetect
    d = torch.from_numpy(pm).cuda()
    Z = pm.shape[0]
    ranges = multi_gpu.partition_layers(Z, world) if cuts is None else list(zip([0] + cuts, cuts + [Z]))
    slabs = [d[z0:z1].contiguous() for z0, z1 in ranges]
    got, st = multi_gpu.voxel2obj_global(slabs, ranges, Z, r, sigma, off, buf, thd,
                                         coll=multi_gpu._LocalCollectives(len(ranges)), return_stats=True)
    want, st1 = fplobjdetect.voxel2obj_device(d, r, sigma, off, buf, thd, return_stats=True)
    return got, st, want, st1


@pytest.mark.parametrize("shape,kind,world,r,sigma,buf,thd", [
    ((96, 70, 80), "blobs", 2, 8, 2.0, 5, 0),
    ((120, 64, 72), "uniform", 3, 6, 1.5, 0, 0),
    ((90, 60, 60), "ties", 4, 5, 2.5, (2, 3, 4), 0),
    ((150, 90, 100), "blobs", 3, 27, 5.0, 30, 0),
    ((80, 50, 50), "uniform", 2, 7, 0.0, 3, 0.6),
    ((64, 48, 48), "zeros", 2, 4, 1.0, 0, 0),
    ((64, 48, 48), "saturated", 2, 4, 1.0, 0, 0),
])
def test_global_equals_single(shape, kind, world, r, sigma, buf, thd):
    pm = cases.prob_map(shape, 41, kind, peaks_per_50cube=12.0)
    got, st, want, st1 = _run(pm, world, r, sigma, (3, -7, 11), buf, thd)
    if want["conf"].size:
        assert st["threshold"] == st1["threshold"]
    assert np.array_equal(got["locs"], want["locs"])
    assert np.array_equal(got["conf"], want["conf"])


def test_global_uneven_and_thin_slabs():
    """Slabs thinner than the halo (a rank's extended slab spans several neighbours) and very uneven cuts."""
    pm = cases.prob_map((100, 56, 60), 43, "blobs", peaks_per_50cube=20.0)
    got, st, want, _ = _run(pm, None, 9, 2.0, (0, 0, 0), 4, 0, cuts=[5, 11, 40, 44, 90])
    assert want["conf"].size > 10
    assert np.array_equal(got["locs"], want["locs"]) and np.array_equal(got["conf"], want["conf"])


def test_global_with_nan_is_empty():
    pm = cases.prob_map((60, 40, 40), 3, "blobs")
    pm[30, 6, 7] = np.nan
    got, st, want, _ = _run(pm, 2, 4, 1.0, (0, 0, 0), 0, 0)
    assert got["conf"].size == 0 and want["conf"].size == 0
