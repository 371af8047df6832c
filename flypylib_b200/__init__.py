"""flypylib_b200 -- B200-native (sm_100a) implementation of flypylib's T-bar detection hot path.

Drop-in surface (same names / arguments / results as janelia-flyem/flypylib):
    fplmodels.vgg_like / vgg_like2 / unet_like2 (+ baseline_model, unet_like, unet_like3/4/4b)
    fplnetwork.FplNetwork (infer, train, make_*_parallel, save_network), load_network
    fplobjdetect.voxel2obj, full_roi_inference, evaluate_substacks, obj_pr*, gen_batches
    fplsynapses json formats
All array work runs in hand-written CUDA kernels behind the C ABI of include/fpl_b200.h
(libfplb200.so); there is no CPU fallback.
"""
from . import fplutils  # noqa: F401

__all__ = ["fplutils", "fplobjdetect", "fplmodels", "fplnetwork", "fplsynapses", "multi_gpu"]


def __getattr__(name):
    if name in ("fplobjdetect", "fplmodels", "fplnetwork", "fplsynapses", "multi_gpu", "FplNetwork"):
        import importlib
        if name == "FplNetwork":
            return importlib.import_module(".fplnetwork", __name__).FplNetwork
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
