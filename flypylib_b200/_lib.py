"""ctypes binding of libfplb200.so (the C ABI declared in include/fpl_b200.h).

There is deliberately no fallback: if the shared library is missing, or no sm_100 GPU is
visible, every compute entry point raises.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfplb200.so")

FPL_OK = 0
PREC_FP32, PREC_BF16, PREC_TF32 = 0, 1, 2
ARCH_VGG_LIKE, ARCH_VGG_LIKE2, ARCH_UNET_LIKE2 = 1, 2, 3
ARCH_BASELINE, ARCH_UNET_LIKE, ARCH_UNET_LIKE3, ARCH_UNET_LIKE4, ARCH_UNET_LIKE4B = 4, 5, 6, 7, 8
ARCH_RESNET_LIKE = 9
ARCH_UNET_LIKE_VOL = 10


class FplError(RuntimeError):
    pass


class V2OParams(ctypes.Structure):
    _fields_ = [
        ("obj_min_dist", ctypes.c_int32),
        ("lw", ctypes.c_int32),
        ("h_weights", ctypes.POINTER(ctypes.c_double)),
        ("thd", ctypes.c_double),
        ("rank_lo", ctypes.c_int64),
        ("rank_hi", ctypes.c_int64),
        ("gamma", ctypes.c_float),
        ("buffer_xyz", ctypes.c_int32 * 3),
        ("offset_xyz", ctypes.c_double * 3),
    ]


_lib = None
_lock = threading.Lock()

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
vp = ctypes.c_void_p

_SIGNATURES = {
    "fpl_version": (ctypes.c_int, []),
    "fpl_last_error": (ctypes.c_char_p, []),
    "fpl_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "fpl_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(vp)]),
    "fpl_ctx_destroy": (ctypes.c_int, [vp]),
    "fpl_ctx_workspace_bytes": (ctypes.c_int, [vp, c_i64p]),
    "fpl_ctx_launch_count": (ctypes.c_int, [vp, c_i64p]),
    "fpl_ctx_release_workspace": (ctypes.c_int, [vp]),
    "fpl_ctx_profile_begin": (ctypes.c_int, [vp]),
    "fpl_ctx_profile_end": (ctypes.c_int, [vp, c_f64p, c_f64p, c_i64p]),
    "fpl_v2o_smooth": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.POINTER(V2OParams), vp, vp]),
    "fpl_v2o_threshold": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.POINTER(V2OParams), c_f64p, vp]),
    "fpl_v2o_detect": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.POINTER(V2OParams), ctypes.c_double, vp, ctypes.c_int64,
                                      c_i64p, c_i64p, vp]),
    "fpl_voxel2obj": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.POINTER(V2OParams), vp, ctypes.c_int64, c_i64p, c_f64p,
                                     c_i64p, vp]),
    "fpl_v2o_detect_seg": (ctypes.c_int, [vp, vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.POINTER(V2OParams), ctypes.c_int32, ctypes.c_int32,
                                          vp, ctypes.c_int64, vp, vp]),
    "fpl_v2o_hist_level": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                          ctypes.c_int, vp, vp, vp]),
    "fpl_v2o_slab_begin": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.POINTER(V2OParams), ctypes.c_double, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(vp), c_i64p, vp]),
    "fpl_v2o_slab_round": (ctypes.c_int, [vp, vp, ctypes.c_int64, c_i64p, c_i64p, vp]),
    "fpl_v2o_slab_suppress": (ctypes.c_int, [vp, vp, ctypes.c_int64, vp]),
    "fpl_v2o_slab_round_pack": (ctypes.c_int, [vp, vp, ctypes.c_int64, ctypes.c_int64, vp]),
    "fpl_v2o_slab_apply_blocks": (ctypes.c_int, [vp, vp, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, vp]),
    "fpl_v2o_slab_end": (ctypes.c_int, [vp, vp, ctypes.c_int64, c_i64p, c_i64p, vp]),
    "fpl_net_create": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(vp)]),
    "fpl_net_destroy": (ctypes.c_int, [vp]),
    "fpl_net_info": (ctypes.c_int, [vp, c_i32p, c_i32p, c_i32p, c_i32p]),
    "fpl_net_num_weights": (ctypes.c_int, [vp, c_i32p]),
    "fpl_net_weight_size": (ctypes.c_int, [vp, ctypes.c_int32, c_i64p]),
    "fpl_net_set_weights": (ctypes.c_int, [vp, ctypes.POINTER(vp), ctypes.c_int32, ctypes.c_int]),
    "fpl_net_set_tile_multiplier": (ctypes.c_int, [vp, ctypes.c_int32]),
    "fpl_net_out_size": (ctypes.c_int, [vp, ctypes.c_int32, c_i32p]),
    "fpl_net_forward_tiles": (ctypes.c_int, [vp, vp, ctypes.c_int32, ctypes.c_int32, vp, vp]),
    "fpl_train_create": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]),
    "fpl_train_destroy": (ctypes.c_int, [vp]),
    "fpl_train_set_precision": (ctypes.c_int, [vp, ctypes.c_int]),
    "fpl_train_sizes": (ctypes.c_int, [vp, c_i64p, c_i64p]),
    "fpl_train_forward_backward": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, ctypes.c_float, ctypes.c_uint64,
                                                  c_f64p, c_i64p, vp]),
    "fpl_train_apply": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_float, ctypes.c_float, vp]),
    "fpl_net_infer_volume": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_int32, ctypes.c_int32, vp, vp]),
    "fpl_net_infer_slab": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                          ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int64, vp, vp]),
}


def exported_symbols():
    """Names every entry point include/fpl_b200.h declares (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def lib():
    """Load libfplb200.so once.  Raises FplError when it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise FplError(
                    "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or make -C flypylib_b200/csrc).  flypylib_b200 has no CPU fallback." % LIB_PATH)
            l = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(l, name)        # AttributeError here = stale build of the library
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def check(rc, what=""):
    if rc != FPL_OK:
        msg = lib().fpl_last_error().decode("utf-8", "replace")
        raise FplError("%s failed (code %d): %s" % (what or "libfplb200 call", rc, msg))


_contexts = {}


class Context:
    """One fpl_ctx per device, shared by every caller in the process."""

    def __init__(self, device):
        self.device = int(device)
        h = vp()
        check(lib().fpl_ctx_create(self.device, ctypes.byref(h)), "fpl_ctx_create")
        self.handle = h

    def launch_count(self):
        v = ctypes.c_int64()
        check(lib().fpl_ctx_launch_count(self.handle, ctypes.byref(v)))
        return v.value

    def workspace_bytes(self):
        v = ctypes.c_int64()
        check(lib().fpl_ctx_workspace_bytes(self.handle, ctypes.byref(v)))
        return v.value

    def release_workspace(self):
        check(lib().fpl_ctx_release_workspace(self.handle), "fpl_ctx_release_workspace")

    def profile_begin(self):
        check(lib().fpl_ctx_profile_begin(self.handle), "fpl_ctx_profile_begin")

    def profile_end(self):
        """-> {family: (ms, work, count)}"""
        ms = (ctypes.c_double * 8)(); work = (ctypes.c_double * 8)(); cnt = (ctypes.c_int64 * 8)()
        check(lib().fpl_ctx_profile_end(self.handle, ms, work, cnt), "fpl_ctx_profile_end")
        names = ["conv3", "conv1", "first", "netaux", "gauss", "select", "nms", "tiler"]
        return {n: (ms[i], work[i], cnt[i]) for i, n in enumerate(names)}

    def close(self):
        if self.handle:
            lib().fpl_ctx_destroy(self.handle)
            self.handle = None


def context(device=None):
    import torch
    if device is None:
        if not torch.cuda.is_available():
            raise FplError("no CUDA device visible: flypylib_b200 runs only on B200 (sm_100a); "
                           "there is no CPU fallback")
        device = torch.cuda.current_device()
    device = int(device)
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def current_stream_ptr(device):
    import torch
    return vp(torch.cuda.current_stream(device).cuda_stream)
