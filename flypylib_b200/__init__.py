"""flypylib_b200 -- B200-native (sm_100a) implementation of flypylib's T-bar detection hot path.

Drop-in surface (same names / arguments / results as janelia-flyem/flypylib):
    fplmodels.vgg_like / vgg_like2 / unet_like2
    fplnetwork.FplNetwork (infer, make_infer_parallel, ...)
    fplobjdetect.voxel2obj
All array work runs in hand-written CUDA kernels behind the C ABI of include/fpl_b200.h
(libfplb200.so); there is no CPU fallback.
"""
from . import fplutils  # noqa: F401

__all__ = ["fplutils", "fplobjdetect", "fplmodels", "fplnetwork", "multi_gpu"]


def __getattr__(name):
    if name in ("fplobjdetect", "fplmodels", "fplnetwork", "multi_gpu", "FplNetwork"):
        import importlib
        if name == "FplNetwork":
            return importlib.import_module(".fplnetwork", __name__).FplNetwork
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
