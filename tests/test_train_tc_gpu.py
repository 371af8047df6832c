"""GPU: the three tensor-core contractions of the training step (csrc/train_tc.cuh: forward, dgrad, wgrad as tcgen05
GEMMs on gathered float32 tensors) one by one against torch float64 conv3d / autograd, at the layer shapes of the VGG
builders (flypylib/fplmodels.py:102-172).  Tolerances, relative to the largest entry of the result: 'tf32' (bf16 hi/lo
split operands, three contractions, fp32 accumulation) 2e-5; 'bf16' (one contraction) 2e-2."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PREC = {"bf16": 1, "tf32": 2}
TOL = {"bf16": 2e-2, "tf32": 2e-5}
# (n, din, k, cin, cout)
SHAPES = [(3, 12, 3, 1, 48), (3, 10, 3, 48, 48), (5, 7, 1, 48, 96), (4, 6, 1, 96, 96), (7, 3, 3, 48, 48),
          (2, 22, 3, 48, 48)]


def _call(what, prec, act, w, out, scratch, n, din, k, cin, cout):
    import torch
    from flypylib_b200 import _lib
    lib = _lib.lib()
    ctx = _lib.context()
    vp = ctypes.c_void_p
    lib.fpl_debug_train_tc.restype = ctypes.c_int
    lib.fpl_debug_train_tc.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp] + [ctypes.c_int] * 5 + [vp]
    _lib.check(lib.fpl_debug_train_tc(ctx.handle, what, prec, act.data_ptr(), w.data_ptr(), out.data_ptr(),
                                      scratch.data_ptr() if scratch is not None else None, n, din, k, cin, cout,
                                      _lib.current_stream_ptr(torch.cuda.current_device())), "fpl_debug_train_tc")
    torch.cuda.synchronize()


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("shape", SHAPES)
def test_contractions_vs_torch_float64(shape, precision):
    import torch
    import torch.nn.functional as F
    n, din, k, cin, cout = shape
    dout = din - (k - 1)
    g = torch.Generator().manual_seed(1000 + din * 7 + cin)
    x = torch.randn((n, din, din, din, cin), generator=g, dtype=torch.float32)
    w = torch.randn((k, k, k, cin, cout), generator=g, dtype=torch.float32) * 0.2
    dy = torch.randn((n, dout, dout, dout, cout), generator=g, dtype=torch.float32)
    x64 = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    w64 = w.double().permute(4, 3, 0, 1, 2).clone().requires_grad_(True)
    y64 = F.conv3d(x64, w64)
    y64.backward(dy.double().permute(0, 4, 1, 2, 3))
    want_y = y64.detach().permute(0, 2, 3, 4, 1).numpy()
    want_dx = x64.grad.permute(0, 2, 3, 4, 1).numpy()
    want_dw = w64.grad.permute(2, 3, 4, 1, 0).numpy()
    xd, wd, dyd = x.cuda(), w.cuda(), dy.cuda()
    tol = TOL[precision]
    p = PREC[precision]

    def check(name, got, want):
        err = float(np.abs(got.cpu().numpy().astype(np.float64) - want).max())
        scale = float(np.abs(want).max())
        print("%s %s %s: max|err| %.3g of scale %.3g" % (name, shape, precision, err, scale))
        assert err < tol * scale, "%s: %g vs scale %g" % (name, err, scale)

    y = torch.full((n, dout, dout, dout, cout), float("nan"), device="cuda")
    _call(0, p, xd, wd, y, None, n, din, k, cin, cout)
    check("forward", y, want_y)
    dw = torch.full((k, k, k, cin, cout), float("nan"), device="cuda")
    _call(2, p, xd, dyd, dw, None, n, din, k, cin, cout)
    check("wgrad", dw, want_dw)
    if cin > 1:
        dp = dout + 2 * (k - 1)
        scratch = torch.empty((n * dp ** 3 * cout,), device="cuda")
        dx = torch.full((n, din, din, din, cin), float("nan"), device="cuda")
        _call(1, p, dyd, wd, dx, scratch, n, din, k, cin, cout)
        check("dgrad", dx, want_dx)
