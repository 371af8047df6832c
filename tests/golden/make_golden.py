"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python -m tests.golden.make_golden            (needs /root/reference; container only)

Outputs (committed):
  tests/golden/voxel2obj_golden.npz   reference ``voxel2obj`` outputs for every case in
                                      ``cases.VOXEL2OBJ_CASES`` + sha256 of the smoothed padded map
                                      and the threshold, obtained with the very calls the reference
                                      makes (fplobjdetect.py:159-183) on SciPy/NumPy of this image
  tests/golden/infer_tiler_golden.npz reference ``FplNetwork.infer`` tiling/scatter run unmodified
                                      with a deterministic fake ``infer_network`` (fplnetwork.py:136-189)
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from tests.golden import cases  # noqa: E402


def golden_voxel2obj(ref):
    from scipy import ndimage
    out = {}
    for name, shape, seed, kind, r, sigma, thd, buf, off in cases.VOXEL2OBJ_CASES:
        pred = cases.prob_map(shape, seed, kind)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = ref.fplobjdetect.voxel2obj(pred.copy(), r, sigma, off, buf, thd)
            # the same calls the reference makes, to pin the intermediate stages
            p = np.pad(pred, r, "constant")
            s = ndimage.gaussian_filter(p, sigma, truncate=2.0)
        if r > 0:
            s[:r] = 0; s[:, :r] = 0; s[:, :, :r] = 0; s[-r:] = 0; s[:, -r:] = 0; s[:, :, -r:] = 0
        else:
            s[...] = 0
        t = np.maximum(np.percentile(s, 97), thd)
        out[name + "/locs"] = res["locs"]
        out[name + "/conf"] = res["conf"]
        out[name + "/thresh"] = np.asarray(t)
        out[name + "/smooth_sha256"] = np.frombuffer(
            hashlib.sha256(np.ascontiguousarray(s).tobytes()).digest(), dtype=np.uint8)
        print("%-26s dets=%5d thresh=%r" % (name, res["conf"].size, t))
    np.savez_compressed(os.path.join(HERE, "voxel2obj_golden.npz"), **out)


_FakeNet = cases.FakeNet


def golden_infer_tiler(ref):
    out = {}
    specs = [
        ("vgg_like_37x41x45", (37, 41, 45), (22, 22, 22), (7, 7, 7), (4, 4, 4), 1),
        ("vgg2_like_50_ngpu4", (50, 50, 50), (28, 28, 28), (10, 10, 10), (4, 4, 4), 4),
        ("unet_like_61x40x33", (61, 40, 33), (30, 30, 30), (9, 9, 9), (1, 1, 1), 3),
        ("small_than_tile", (20, 25, 30), (30, 30, 30), (9, 9, 9), (1, 1, 1), 1),
    ]
    for name, shape, infer_sz, off, stride, n_gpu in specs:
        import zlib; rng = np.random.default_rng(zlib.crc32(name.encode()))
        img = rng.standard_normal(shape).astype(np.float32)
        net = ref.fplnetwork.FplNetwork.__new__(ref.fplnetwork.FplNetwork)
        net.infer_sz, net.rf_offset, net.rf_stride, net.n_gpu = infer_sz, off, stride, n_gpu
        net.rf_size = tuple(2 * o + s for o, s in zip(off, stride))
        net.infer_network = _FakeNet(infer_sz, off, stride)
        pred = net.infer(img)
        out[name + "/image"] = img
        out[name + "/pred"] = pred
        out[name + "/spec"] = np.asarray(list(shape) + list(infer_sz) + list(off) + list(stride) + [n_gpu])
        out[name + "/n_batch"] = np.asarray(net.infer_network.calls[0][0][0])
        print("%-22s pred %s batch %s" % (name, pred.shape, net.infer_network.calls[0]))
    np.savez_compressed(os.path.join(HERE, "infer_tiler_golden.npz"), **out)


if __name__ == "__main__":
    ref = ref_loader.load()
    golden_voxel2obj(ref)
    golden_infer_tiler(ref)
