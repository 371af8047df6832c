"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python -m tests.golden.make_golden            (needs /root/reference; container only)

Outputs (committed):
  tests/golden/voxel2obj_golden.npz   reference ``voxel2obj`` outputs for every case in
                                      ``cases.VOXEL2OBJ_CASES`` + sha256 of the smoothed padded map
                                      and the threshold, obtained with the very calls the reference
                                      makes (fplobjdetect.py:159-183) on SciPy/NumPy of this image
  tests/golden/voxel2obj_seg_golden.npz  reference ``voxel2obj`` with seg / seg_dilate / seg_sz_thd / seg_force
                                      (fplobjdetect.py:161-224) run unmodified on ``cases.VOXEL2OBJ_SEG_CASES``
  tests/golden/infer_tiler_golden.npz reference ``FplNetwork.infer`` tiling/scatter run unmodified
                                      with a deterministic fake ``infer_network`` (fplnetwork.py:136-189)
  tests/golden/substack_golden.npz    reference ``fri_get_image`` (fplobjdetect.py:1021-1112) run unmodified on an
                                      in-memory store: substack cubes, normalisation, norm .txt lines
  tests/golden/eval_golden.json       reference ``obj_pr`` / ``obj_pr_curve`` / ``aggregate_pr``
                                      (fplobjdetect.py:320-453) run unmodified with the exhaustive
                                      ``oracle/match_oracle.py`` standing in for the PuLP solver, and the
                                      reference json writers/reader (fplsynapses.py:11-111) run unmodified
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


def golden_voxel2obj_seg(ref):
    """reference voxel2obj run unmodified with seg / seg_dilate / seg_sz_thd / seg_force (fplobjdetect.py:161-224)"""
    out = {}
    for name, shape, seed, kind, r, sigma, thd, buf, off, dil, szt, force in cases.VOXEL2OBJ_SEG_CASES:
        pred = cases.prob_map(shape, seed, kind)
        seg = cases.segmentation(shape, seed + 100)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = ref.fplobjdetect.voxel2obj(pred.copy(), r, sigma, off, buf, thd, seg=seg.copy(), seg_dilate=dil,
                                             seg_sz_thd=szt, seg_force=force)
            plain = ref.fplobjdetect.voxel2obj(pred.copy(), r, sigma, off, buf, thd)
        out[name + "/locs"] = res["locs"]
        out[name + "/conf"] = res["conf"]
        print("%-26s dets=%5d (without seg: %d)" % (name, res["conf"].size, plain["conf"].size))
    np.savez_compressed(os.path.join(HERE, "voxel2obj_seg_golden.npz"), **out)


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


def eval_cases():
    """Seeded (pred, gt) detection sets for the scoring goldens: jittered ground truth + spurious
    predictions + misses, continuous coordinates (no exact cost ties)."""
    out = []
    for seed, n_gt, n_extra, n_miss, box in [(1, 12, 4, 2, 120.0), (2, 9, 0, 0, 60.0), (3, 14, 6, 5, 80.0),
                                             (4, 6, 10, 1, 50.0)]:
        rng = np.random.default_rng(seed)
        gt = rng.uniform(0, box, (n_gt, 3))
        keep = np.ones(n_gt, bool); keep[rng.choice(n_gt, n_miss, replace=False)] = False
        pred = np.concatenate([gt[keep] + rng.normal(0, 6.0, (keep.sum(), 3)), rng.uniform(0, box, (n_extra, 3))])
        conf = rng.uniform(0.2, 1.0, pred.shape[0])
        out.append((pred, conf, gt))
    return out


def golden_eval(ref):
    import json
    import tempfile
    from oracle import match_oracle
    import flypylib.fplsynapses as ref_syn          # unmodified reference module (stubs for dvid/h5py)
    F = ref.fplobjdetect
    F.obj_match = match_oracle.obj_match            # PuLP stand-in: exact exhaustive solver
    thresholds = np.array([0.0, 0.3, 0.5, 0.7, 0.9, 2.0])
    doc = {"thresholds": thresholds.tolist(), "pr": [], "formats": []}
    curves = []
    for pred, conf, gt in eval_cases():
        for allow_mult in (False, True):
            r = F.obj_pr(pred, gt, 27.0, allow_mult=allow_mult)
            c = F.obj_pr_curve({"locs": pred, "conf": conf}, {"locs": gt}, 27.0, thresholds, allow_mult=allow_mult)
            if not allow_mult:
                curves.append(c)
            doc["pr"].append({"pred": pred.tolist(), "conf": conf.tolist(), "gt": gt.tolist(), "allow_mult": allow_mult,
                              "num_tp": int(r.num_tp), "tot_pred": int(r.tot_pred), "tot_gt": int(r.tot_gt),
                              "pp": float(r.pp), "rr": float(r.rr),
                              "match_cost": float((np.sqrt(((pred[:, None] - gt[None]) ** 2).sum(2)) - 27.0)[r.match].sum()),
                              "curve": {k: np.asarray(getattr(c, k)).tolist() for k in ("num_tp", "tot_pred", "tot_gt", "pp", "rr")}})
    agg = F.aggregate_pr(curves)
    doc["aggregate"] = {k: np.asarray(getattr(agg, k)).tolist() for k in ("num_tp", "tot_pred", "tot_gt", "pp", "rr")}
    e = F.obj_pr(np.zeros((0, 3)), np.zeros((4, 3)), 27.0)
    doc["empty_pred"] = [int(e.num_tp), int(e.tot_pred), int(e.tot_gt), e.pp, e.rr, e.match]
    e = F.obj_pr(np.zeros((3, 3)), np.zeros((0, 3)), 27.0)
    doc["empty_gt"] = [int(e.num_tp), int(e.tot_pred), int(e.tot_gt), e.pp, e.rr, e.match]
    # wire formats
    rng = np.random.default_rng(7)
    tb = {"locs": rng.uniform(0, 500, (9, 3)), "conf": rng.uniform(0, 1, 9)}
    tb["conf"][0] = 0.0005; tb["conf"][1] = 0.9995
    dvid = ref_syn.tbars_to_json_format(tb, labels=np.arange(9) * 7)
    dvid_nolab = ref_syn.tbars_to_json_format(tb, user_name="someone")
    rav = ref_syn.tbars_to_json_format_raveler(tb)
    mixed = dvid_nolab + [{"Kind": "PostSyn", "Pos": [1, 2, 3], "Prop": {}},
                          {"Kind": "PreSyn", "Pos": [4, 5, 6], "Prop": {"err": "0.25"}}]
    back_dvid = ref_syn.load_from_json(json.dumps(mixed))
    back_nested = ref_syn.load_from_json(json.dumps([dvid_nolab]))
    back_rav = ref_syn.load_from_json(json.dumps(rav))
    back_buf = ref_syn.load_from_json(json.dumps(dvid_nolab), vol_sz=500, buffer=(60, 40, 80))
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        path = f.name
    ref_syn.tbars_to_json_format(tb, json_file=path)
    file_text = open(path).read(); os.unlink(path)

    def pack(t):
        return {k: [None if (isinstance(x, float) and x != x) else x for x in np.asarray(v).tolist()] if np.asarray(v).ndim == 1
                else np.asarray(v).tolist() for k, v in t.items()}
    doc["formats"] = {"locs": tb["locs"].tolist(), "conf": tb["conf"].tolist(), "dvid_labels": dvid, "dvid_user": dvid_nolab,
                      "raveler": rav, "mixed": mixed, "back_dvid": pack(back_dvid), "back_nested": pack(back_nested),
                      "back_raveler": pack(back_rav), "back_buffer": pack(back_buf), "file_text": file_text}
    json.dump(doc, open(os.path.join(HERE, "eval_golden.json"), "w"))
    print("eval goldens: %d pr cases" % len(doc["pr"]))


class FakeStore(object):
    """Array-like volume store for the reference's "n5" branch of fri_get_image (shape + slicing)."""

    def __init__(self, a):
        self.a, self.shape = a, a.shape

    def __getitem__(self, k):
        return self.a[k]


def golden_substack_images(ref):
    """Reference fri_get_image (fplobjdetect.py:1021-1112) run unmodified on an in-memory volume store:
    cube extraction with zero padding at the faces, the global_frac normalisation and the norm .txt line."""
    import tempfile
    sys.modules["z5py"].dataset.Dataset = FakeStore         # the reference isinstance()-checks its store type
    F = ref.fplobjdetect
    vol = cases.em_volume((70, 64, 80), seed=9)
    vol[:, :5] = 0; vol[:, -4:] = 255                        # values outside (1,200) for the filtered mean
    out = {"volume": vol}
    specs = [("inside", F.szyx(24, 20, 20, 24), 6, (128., 33.)), ("corner", F.szyx(24, 0, 0, 0), 8, (120., 30., 0.25)),
             ("far", F.szyx(32, 48, 40, 56), 10, (128., 33., 0.0)), ("outside", F.szyx(16, 200, 10, 10), 4, (128., 33.))]
    with tempfile.TemporaryDirectory() as nd:
        for name, ss, buf, norm in specs:
            img, _ = F.fri_get_image([ss, None, None, list(norm), buf, None, nd, "grayscale"], FakeStore(vol), True)
            out[name + "/spec"] = np.asarray([ss.size, ss.z, ss.y, ss.x, buf], dtype=np.int64)
            out[name + "/norm"] = np.asarray(norm, dtype=np.float64)
            out[name + "/none"] = np.asarray(img is None)
            if img is not None:
                out[name + "/image"] = img
                txt = open("%s/%d_%d_%d_%d.txt" % (nd, ss.size, ss.z, ss.y, ss.x)).read()
                out[name + "/txt"] = np.frombuffer(txt.encode(), dtype=np.uint8)
            print("%-8s %s" % (name, None if img is None else (img.shape, img.dtype, float(img.mean()))))
    np.savez_compressed(os.path.join(HERE, "substack_golden.npz"), **out)


def gen_batches_volumes():
    """Two small training volumes (image float32, labels {0,1}, mask {0,1}) for the gen_batches goldens."""
    vols = []
    for seed in (31, 32):
        rng = np.random.default_rng(seed)
        shape = (26, 24, 28)
        im = rng.standard_normal(shape).astype(np.float32)
        ll = (rng.random(shape) < 0.08).astype(np.uint8)
        mm = (rng.random(shape) < 0.85).astype(np.uint8)
        vols.append((im, ll, mm))
    return vols


def golden_gen_batches(ref):
    """Reference gen_batches (fplobjdetect.py:27-130) run unmodified with an in-memory stand-in for h5py.File."""
    vols = gen_batches_volumes()
    store = {}
    for i, (im, ll, mm) in enumerate(vols):
        store["im%d.h5" % i] = im; store["p%d_labels.h5" % i] = ll; store["p%d_mask.h5" % i] = mm

    class FakeFile(object):
        def __init__(self, name, mode="r"):
            self.name = name

        def __getitem__(self, key):
            return store[self.name].copy()

    ref.fplobjdetect.h5py.File = FakeFile
    out = {}
    train = tuple(("im%d.h5" % i, "p%d_" % i) for i in range(len(vols)))
    for tag, ctx, bs, is_mask in [("cls", (8, 8, 8), 6, False), ("mask", (10, 10, 10), 4, True)]:
        np.random.seed(1234)
        g = ref.fplobjdetect.gen_batches(train, ctx, bs, is_mask=is_mask)
        for k in range(3):
            d, l = next(g)
            out["%s/data%d" % (tag, k)] = d.copy(); out["%s/labels%d" % (tag, k)] = l.copy()
    np.savez_compressed(os.path.join(HERE, "gen_batches_golden.npz"), **out)
    print("gen_batches goldens:", sorted(out)[:4], "...")


if __name__ == "__main__":
    ref = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == "seg":           # only the segmentation-aware goldens
        golden_voxel2obj_seg(ref)
        sys.exit(0)
    golden_voxel2obj(ref)
    golden_voxel2obj_seg(ref)
    golden_infer_tiler(ref)
    golden_eval(ref)
    golden_substack_images(ref)
    golden_gen_batches(ref)
