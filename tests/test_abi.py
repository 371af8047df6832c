"""CPU: the C-ABI library loads and exports every symbol include/fpl_b200.h declares; the product
refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import flypylib_b200
from flypylib_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "fpl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(_lib.exported_symbols()) == set(names)


def test_version_and_error_string():
    l = _lib.lib()
    assert l.fpl_version() == 1
    assert isinstance(l.fpl_last_error(), bytes)


def test_params_struct_layout_matches_header():
    # offsets implied by the C declaration (natural alignment)
    assert _lib.V2OParams.obj_min_dist.offset == 0
    assert _lib.V2OParams.lw.offset == 4
    assert _lib.V2OParams.h_weights.offset == 8
    assert _lib.V2OParams.thd.offset == 16
    assert _lib.V2OParams.rank_lo.offset == 24
    assert _lib.V2OParams.rank_hi.offset == 32
    assert _lib.V2OParams.gamma.offset == 40
    assert _lib.V2OParams.buffer_xyz.offset == 44
    assert _lib.V2OParams.offset_xyz.offset == 56
    assert ctypes.sizeof(_lib.V2OParams) == 80


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flypylib_b200 import fplobjdetect
    with pytest.raises(_lib.FplError):
        fplobjdetect.voxel2obj(np.zeros((8, 8, 8), np.float32), 2, 1.0)
    h = ctypes.c_void_p()
    rc = _lib.lib().fpl_ctx_create(0, ctypes.byref(h))
    assert rc == -4  # FPL_ENODEV


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flypylib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_save_network_load_network_roundtrip(tmp_path):
    """fplnetwork.save_network / load_network (reference fplnetwork.py:32-44, 81-97): the pickled network and
    its weight side file restore builder, receptive-field info, compile_args and weights (host-only, no GPU)."""
    import numpy as np
    from flypylib_b200 import fplmodels, fplnetwork
    net = fplnetwork.FplNetwork(fplmodels.unet_like2)
    rng = np.random.default_rng(3)
    w = [rng.standard_normal(a.shape).astype(np.float32) for a in net.train_single.get_weights()]
    net.train_single.set_weights(w)
    net.set_precision("tf32")
    net._set_infer()
    net.tile_multiplier = 1
    path = str(tmp_path / "net.p")
    net.save_network(path)
    assert net.infer_network is not None                      # the live network is untouched
    back = fplnetwork.load_network(path)
    assert back.rf_size == net.rf_size and back.rf_offset == net.rf_offset and back.infer_sz == net.infer_sz
    assert back.compile_args.keys() == net.compile_args.keys()
    assert back.train_single.precision == "tf32" and back.infer_network.precision == "tf32"
    for a, b in zip(back.infer_network.get_weights(), w):
        assert np.array_equal(a, b)


def test_error_behaviour_of_the_python_shims():
    """Errors raised before any GPU work, as the reference raises them: un-built network (fplnetwork.py:141-142),
    wrong input rank / dtype, out-of-scope branches."""
    import numpy as np
    import pytest
    from flypylib_b200 import fplmodels, fplnetwork, fplobjdetect
    net = fplnetwork.FplNetwork(fplmodels.vgg_like)
    with pytest.raises(AssertionError, match="network has not been trained"):
        net.infer(np.zeros((8, 8, 8), np.float32))
    with pytest.raises((OSError, AssertionError)):
        net.infer("no_such_volume.h5")
    with pytest.raises(ValueError):
        fplobjdetect.voxel2obj(np.zeros((4, 4), np.float32), 3, 1.0)
    with pytest.raises(TypeError):
        fplobjdetect.voxel2obj(np.zeros((4, 4, 4), np.float64), 3, 1.0)
    from flypylib_b200._lib import FplError
    with pytest.raises((FplError, ValueError)):          # no GPU here: no CPU fallback; on a GPU box: shape mismatch
        fplobjdetect.voxel2obj(np.zeros((4, 4, 4), np.float32), 3, 1.0, seg=np.zeros((4, 4, 5)))
    with pytest.raises(ValueError):
        fplmodels.vgg_like()[0].set_weights([np.zeros(3, np.float32)])
    m = fplmodels.vgg_like2()[0]
    assert m.count_params() == 265777                      # 264 048 conv weights + BN + final bias
    assert fplmodels.unet_like2()[1] == (24, 9, 1) and fplmodels.vgg_like()[2] == 102
