"""GPU parity of the CNN half: CUDA network (through the C ABI) against the torch-CPU float64
restatement of the Keras graphs (oracle/models_oracle.py; parity unpinned, see its header) and
against the reference tiling semantics (pinned by tests/golden/infer_tiler_golden.npz)."""
import numpy as np
import pytest

from oracle import models_oracle as M
from tests.golden import cases

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-3      # north_star: prob maps within 2e-3 max-abs on the fp32/TF32 path
SMALL = {"vgg_like": 30, "vgg_like2": 36, "unet_like2": 36, "baseline_model": 30, "unet_like": 30, "unet_like3": 44,
         "unet_like4": 52, "unet_like4b": 52, "resnet_like": 34, "unet_like_vol": 34}


def _builder(arch):
    from flypylib_b200 import fplmodels
    return getattr(fplmodels, arch)


@pytest.mark.parametrize("arch", ["vgg_like", "vgg_like2", "unet_like2", "baseline_model", "unet_like", "unet_like3",
                                  "unet_like4", "unet_like4b", "resnet_like", "unet_like_vol"])
def test_forward_tiles_fp32_vs_float64_oracle(arch):
    s = SMALL[arch]
    model, rf, infer_sz, _ = _builder(arch)(s)
    model.upsample_output = True
    model.set_precision("fp32")
    w = M.random_weights(arch, seed=4321)
    model.set_weights(w)
    x = np.random.default_rng(7).standard_normal((3, s, s, s)).astype(np.float32)
    got = model.predict(x[..., None], batch_size=2)[..., 0]
    want = M.forward(arch, w, x)
    assert got.shape == want.shape
    err = np.abs(got.astype(np.float64) - want).max()
    assert err < 1e-5, err          # fp32 accumulate vs float64: far inside the 2e-3 budget


@pytest.mark.parametrize("arch", ["vgg_like", "vgg_like2"])
def test_train_graph_predict_is_not_upsampled(arch):
    """The builder's own graph (no UpSampling3D) maps an rf_size patch to one value (fplobjdetect.py:77)."""
    model, rf, _, _ = _builder(arch)(rf_sz := M.ARCHS[arch][1][0])
    model.set_precision("fp32")
    w = M.random_weights(arch, seed=1)
    model.set_weights(w)
    x = np.random.default_rng(3).standard_normal((5, rf_sz, rf_sz, rf_sz)).astype(np.float32)
    got = model.predict(x[..., None], batch_size=5)
    assert got.shape == (5, 1, 1, 1, 1)
    want = M.forward(arch, w, x, upsample=False)
    assert np.abs(got[..., 0] - want).max() < 1e-5


@pytest.mark.parametrize("arch,shape", [("vgg_like", (110, 102, 120)), ("unet_like2", (100, 105, 190)),
                                        ("unet_like3", (100, 120, 100)), ("resnet_like", (120, 102, 110))])
def test_infer_volume_fp32_vs_reference_tiling(arch, shape):
    """FplNetwork.infer (device tiler) == reference tiling (oracle infer_tiler, pinned to the
    reference) driven by the torch restatement of the graph."""
    from flypylib_b200 import fplnetwork
    net = fplnetwork.FplNetwork(_builder(arch))
    w = M.random_weights(arch, seed=99)
    net.train_single.set_weights(w)
    net.set_precision("fp32")
    net._set_infer()
    img = ((cases.em_volume(shape, seed=5).astype(np.float32) - 128.0) / 33.0).astype(np.float32)
    got = net.infer(img)
    assert got.dtype == np.float32 and got.shape == img.shape
    import torch
    ref_net = M.TorchNet(arch, w, dtype=torch.float32)
    want = M.infer_tiler(img, ref_net, net.infer_sz, net.rf_offset, n_gpu=1)
    o = net.rf_offset[0]
    assert np.all(got[:o] == 0) and np.all(got[:, :o] == 0) and np.all(got[:, :, :o] == 0)
    assert np.all(got[-o:] == 0) and np.all(got[:, -o:] == 0) and np.all(got[:, :, -o:] == 0)
    err = np.abs(got - want).max()
    assert err < 1e-4, err
    # uint8-resident path with on-the-fly normalisation gives the same map
    u8 = cases.em_volume(shape, seed=5)
    got2 = net.infer_device(torch.from_numpy(u8).cuda(), normalize=(128.0, 33.0)).cpu().numpy()
    assert np.array_equal(got, got2)
