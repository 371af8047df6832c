"""GPU: one training step of the VGG builders (config 5) against the torch-float64 restatement
(oracle/train_oracle.py; parity unpinned -- Keras/TensorFlow are not installable)."""
import numpy as np
import pytest

from oracle import models_oracle as M
from oracle import train_oracle as T

pytestmark = pytest.mark.gpu


def _setup(arch, batch, seed):
    from flypylib_b200 import fplmodels, fpltrain
    rf = M.ARCHS[arch][1][0]
    model, _, _, _ = getattr(fplmodels, arch)(rf)
    w = M.random_weights(arch, seed=seed)
    model.set_weights(w)
    tr = fpltrain.Trainer(model, rf, batch)
    rng = np.random.default_rng(seed + 1)
    x = rng.standard_normal((batch, rf, rf, rf)).astype(np.float32)
    y = (rng.random(batch) < 0.5).astype(np.uint8)
    return model, tr, w, x, y


@pytest.mark.parametrize("arch,batch", [("vgg_like", 8), ("vgg_like2", 6)])
def test_forward_backward_and_adam_vs_oracle(arch, batch):
    import torch
    model, tr, w, x, y = _setup(arch, batch, 17)
    seed, gb = 12345, batch * 4                      # as if 4 ranks contributed to the global mean
    loss, ok = tr.forward_backward(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), gb, seed)
    want_loss, want_ok, want_g, want_bn = T.forward_backward(arch, w, x, y, gb, seed)
    assert abs(loss - want_loss) < 1e-4 * max(1.0, abs(want_loss))
    assert ok == want_ok
    got_g = T.split_params(arch, tr.grads.cpu().numpy())
    for i, (g, wg) in enumerate(zip(got_g, want_g)):
        scale = max(np.abs(wg).max(), 1e-6)
        assert np.abs(g - wg).max() < 2e-3 * scale + 1e-7, "gradient %d: %g vs scale %g" % (i, np.abs(g - wg).max(), scale)
    got_bn = tr.bn_batch.cpu().numpy()
    o = 0
    for mean, var in want_bn:
        c = mean.size
        assert np.abs(got_bn[o:o + c] - mean).max() < 1e-4
        assert np.abs(got_bn[o + c:o + 2 * c] - var).max() < 1e-4 * max(1.0, var.max())
        o += 2 * c
    # Adam + moving statistics
    tr.apply()
    w64 = [np.asarray(a, dtype=np.float64) for a in w]
    m = [np.zeros_like(a) for a in w64]; v = [np.zeros_like(a) for a in w64]
    want_w, _, _ = T.adam_step(w64, want_g, m, v, want_bn, 1)
    got_w = T.split_params(arch, tr.params.cpu().numpy())
    for i, (a, b) in enumerate(zip(got_w, want_w)):
        # first Adam step moves every trainable weight by ~lr regardless of gradient scale
        assert np.abs(a - b).max() < 2e-4, "weight array %d after Adam: %g" % (i, np.abs(a - b).max())


def test_fplnetwork_train_dropin(tmp_path):
    """FplNetwork.train with a synthetic generator: loss goes down on a learnable toy task, CSV log and
    per-epoch saves are written, the inference network picks up the trained weights."""
    from flypylib_b200 import fplmodels, fplnetwork
    net = fplnetwork.FplNetwork(fplmodels.vgg_like)
    net.make_train_parallel(1, 16, net.rf_size)
    rng = np.random.default_rng(0)

    def gen():
        while True:
            lab = (rng.random(16) < 0.5).astype(np.uint8)
            data = rng.standard_normal((16, 18, 18, 18, 1)).astype(np.float32) * 0.5
            data[lab == 1, 7:11, 7:11, 7:11, 0] += 2.0           # bright blob at the centre = positive
            yield data, lab.reshape(16, 1, 1, 1, 1)

    w0 = net.train_single.get_weights()
    log = tmp_path / "log.csv"
    net.train(gen(), 12, 2, str(log), str(tmp_path / "ckpt"))
    rows = open(log).read().strip().splitlines()
    assert rows[0] == "epoch,acc,loss" and len(rows) == 3
    l0, l1 = float(rows[1].split(",")[2]), float(rows[2].split(",")[2])
    assert l1 < l0
    assert (tmp_path / "ckpt_000.h5.npz").exists() and (tmp_path / "ckpt_001.h5.npz").exists()
    w1 = net.train_single.get_weights()
    assert any(np.abs(a - b).max() > 0 for a, b in zip(w0, w1))
    assert net.infer_network is not None
    assert all(np.array_equal(a, b) for a, b in zip(net.infer_network.get_weights(), w1))
