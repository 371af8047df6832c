"""GPU: the tcgen05 implicit-GEMM path (bf16) against (a) the CUDA-core direct convolution on the same
bf16 C8-blocked tensors and (b) the float64 restatement of the graphs."""
import ctypes

import numpy as np
import pytest

from oracle import models_oracle as M

pytestmark = pytest.mark.gpu


def _predict(arch, s, w, x, precision, force_direct=False):
    from flypylib_b200 import fplmodels, _lib
    lib = _lib.lib()
    lib.fpl_debug_force_direct_conv.argtypes = [ctypes.c_int]
    lib.fpl_debug_force_direct_conv(1 if force_direct else 0)
    try:
        model, _, _, _ = getattr(fplmodels, arch)(s)
        model.upsample_output = True
        model.set_precision(precision)
        model.set_weights(w)
        return model.predict(x[..., None], batch_size=x.shape[0])[..., 0]
    finally:
        lib.fpl_debug_force_direct_conv(0)


@pytest.mark.parametrize("arch,s,n", [("vgg_like2", 36, 2), ("vgg_like2", 52, 3), ("vgg_like", 38, 2),
                                      ("unet_like2", 36, 2), ("vgg_like2", 100, 1)])
def test_bf16_umma_vs_direct_and_oracle(arch, s, n):
    w = M.random_weights(arch, seed=11)
    x = np.random.default_rng(s).standard_normal((n, s, s, s)).astype(np.float32)
    got = _predict(arch, s, w, x, "bf16")
    ref_direct = _predict(arch, s, w, x, "bf16", force_direct=True)
    assert got.shape == ref_direct.shape
    d = np.abs(got - ref_direct).max()
    assert d < 2e-3, "tcgen05 vs direct (same bf16 operands): %g" % d
    if s <= 52:
        want = M.forward(arch, w, x)
        e = np.abs(got.astype(np.float64) - want).max()
        assert e < 2e-2, "bf16 path vs float64 oracle: %g" % e
