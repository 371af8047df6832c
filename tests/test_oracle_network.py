"""CPU: the CNN half of the oracle (oracle/models_oracle.py).

* ``infer_tiler`` against tests/golden/infer_tiler_golden.npz -- outputs of the UNMODIFIED reference
  ``FplNetwork.infer`` (flypylib/fplnetwork.py:136-189) run with ``cases.FakeNet`` as the Keras model
  (tests/golden/make_golden.py:golden_infer_tiler).  This is the pin of the tiling / scatter half.
* ``forward`` (torch float64) against a second, independently written restatement of the same Keras-2
  layer semantics on scipy.ndimage / numpy float64 (no torch, different convolution code).  The network
  arithmetic itself stays "parity unpinned" (Keras/TF are not installable here); this removes the
  single-implementation risk of the restatement, not the missing pin.
"""
import os

import numpy as np
import pytest
from scipy import ndimage

from oracle import models_oracle as M
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "infer_tiler_golden.npz"))
TILER_CASES = sorted({k.split("/")[0] for k in GOLD.files})


@pytest.mark.parametrize("name", TILER_CASES)
def test_infer_tiler_matches_reference_golden(name):
    spec = [int(v) for v in GOLD[name + "/spec"]]
    shape, infer_sz, off, stride, n_gpu = spec[0:3], spec[3:6], spec[6:9], spec[9:12], spec[12]
    img = GOLD[name + "/image"]
    assert list(img.shape) == shape
    net = cases.FakeNet(tuple(infer_sz), tuple(off), tuple(stride))
    pred = M.infer_tiler(img, net, tuple(infer_sz), tuple(off), n_gpu=n_gpu)
    want = GOLD[name + "/pred"]
    assert pred.dtype == want.dtype == np.float32 and pred.shape == want.shape
    assert np.array_equal(pred, want)
    # the staging batch the reference hands to predict(): float64, padded to a multiple of n_gpu (:156-173)
    (bshape, bdtype, bsz), = net.calls
    assert bshape[0] == int(GOLD[name + "/n_batch"]) and bshape[0] % n_gpu == 0
    assert bdtype == "float64" and bsz == n_gpu and tuple(bshape[1:4]) == tuple(infer_sz)
    # the rf_offset-wide border is never written (:178-189)
    o = off
    if all(s > 2 * oo for s, oo in zip(shape, o)):
        inner = np.zeros(shape, bool)
        inner[o[0]:shape[0] - o[0], o[1]:shape[1] - o[1], o[2]:shape[2] - o[2]] = True
        assert not pred[~inner].any()


# ---------------------------------------------------------------------------------------------------
# second restatement: numpy / scipy.ndimage float64, written from the Keras-2 layer definitions
# ---------------------------------------------------------------------------------------------------
def _conv3d_valid(x, kern):
    """x (D,H,W,Cin) float64, kern (k,k,k,Cin,Cout): Keras Conv3D(padding='valid') = cross-correlation."""
    k = kern.shape[0]
    cin, cout = kern.shape[3], kern.shape[4]
    lo, hi = k // 2, k - 1 - k // 2
    out = np.zeros(tuple(s - k + 1 for s in x.shape[:3]) + (cout,))
    for co in range(cout):
        acc = np.zeros(x.shape[:3])
        for ci in range(cin):
            acc += ndimage.correlate(x[..., ci], kern[..., ci, co], mode="constant", cval=0.0)
        out[..., co] = acc[lo:x.shape[0] - hi, lo:x.shape[1] - hi, lo:x.shape[2] - hi]
    return out


def _forward_scipy(arch, weights, x, upsample=True):
    ops, rf, _, final_bias = M.ARCHS[arch]
    t = np.asarray(x, np.float64)[..., None]
    wi, skips = 0, {}
    for op in ops:
        if op[0] in ("C", "CB"):
            kern = np.asarray(weights[wi], np.float64)
            gamma, beta, mean, var = (np.asarray(w, np.float64) for w in weights[wi + 1:wi + 5])
            wi += 5
            t = _conv3d_valid(t, kern)
            t = gamma * (t - mean) / np.sqrt(var + 1e-3) + beta
            if op[0] == "C":
                t = np.maximum(t, 0.0)
        elif op[0] == "CN":
            t = np.maximum(_conv3d_valid(t, np.asarray(weights[wi], np.float64)), 0.0)
            wi += 1
        elif op[0] == "CS":                       # resnet_like shortcut: plain 1x1x1 convolution of a stored tensor
            skips[op[4]] = skips[op[4]] @ np.asarray(weights[wi], np.float64)[0, 0, 0]
            wi += 1
        elif op[0] == "A":
            sk, c = skips[op[1]], op[2]
            t = np.maximum(sk[c:-c, c:-c, c:-c] + t, 0.0)
        elif op[0] == "P":
            d, h, w = (s // 2 for s in t.shape[:3])
            t = t[:2 * d, :2 * h, :2 * w].reshape(d, 2, h, 2, w, 2, -1).max(axis=(1, 3, 5))
        elif op[0] == "S":
            skips[op[1]] = t
        elif op[0] == "U":
            up = t.repeat(2, 0).repeat(2, 1).repeat(2, 2)
            sk, c = skips[op[1]], op[2]
            if c:
                sk = sk[c:-c, c:-c, c:-c]
            t = np.concatenate([up, sk], -1)
        elif op[0] == "F":
            kern = np.asarray(weights[wi], np.float64)
            wi += 1
            t = t @ kern[0, 0, 0]
            if final_bias:
                t = t + float(weights[wi][0])
                wi += 1
            t = 1.0 / (1.0 + np.exp(-t))
    t = t[..., 0]
    if upsample and rf[2] != 1:
        t = t.repeat(rf[2], 0).repeat(rf[2], 1).repeat(rf[2], 2)
    return t


@pytest.mark.parametrize("arch,s", [("vgg_like", 22), ("vgg_like2", 28), ("unet_like2", 28), ("resnet_like", 26),
                                    ("unet_like_vol", 30)])
def test_forward_restatements_agree(arch, s):
    w = M.random_weights(arch, seed=17)
    x = np.random.default_rng(3).standard_normal((s, s, s))
    a = M.forward(arch, w, x[None])[0]
    b = _forward_scipy(arch, w, x)
    assert a.shape == b.shape
    assert np.abs(a - b).max() < 1e-12
