"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (janelia-flyem/flypylib) for golden-vector generation.

/root/reference exists only in the build container, never on the GPU box, so this
module is used solely by ``tests/golden/make_golden.py`` (and by optional
container-only tests).  Nothing under ``flypylib_b200/`` may import it.

The reference cannot be imported as-is (keras / tensorflow / h5py / pulp / diced /
libdvid / z5py / skimage / matplotlib are not installed and there is no network).
Its ``voxel2obj`` (flypylib/fplobjdetect.py:132-257) and ``FplNetwork.infer``
(flypylib/fplnetwork.py:136-189) bodies only need numpy + scipy, so the missing
third-party modules are replaced by inert stubs in ``sys.modules`` and the reference
source files are executed unmodified from where they lie.
"""
import sys
import types
import warnings
from unittest import mock

REFERENCE_ROOT = "/root/reference"

_STUBS = [
    "diced", "libdvid", "libdvid._dvid_python", "z5py", "h5py", "pulp", "skimage",
    "skimage.exposure", "keras", "keras.models", "keras.layers", "keras.layers.core",
    "keras.callbacks", "keras.backend", "tensorflow", "tensorflow.python",
    "tensorflow.python.framework", "tensorflow.python.framework.ops",
    "tensorflow.python.ops", "flyem_syn_eval", "matplotlib", "matplotlib.pyplot",
]


def available():
    import os
    return os.path.isdir(REFERENCE_ROOT + "/flypylib")


def load():
    """Return the reference modules (fplobjdetect, fplnetwork, fplutils)."""
    if not available():
        raise RuntimeError("reference tree %s is not present on this machine" % REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    sys.modules["keras.callbacks"].Callback = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from flypylib import fplobjdetect, fplnetwork, fplutils  # noqa
    return types.SimpleNamespace(fplobjdetect=fplobjdetect, fplnetwork=fplnetwork,
                                 fplutils=fplutils)
