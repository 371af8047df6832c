"""Substack driver (full_roi_inference and fri_* helpers, flypylib/fplobjdetect.py:841-1155).

CPU: cube extraction / normalisation against goldens of the UNMODIFIED reference fri_get_image; control
flow (ROI file, resume pickles, ordering, all.p, rank distribution over gloo) with a fake network and the
oracle voxel2obj standing in for the CUDA one.  GPU: the real pipeline end to end."""
import os
import pickle

import numpy as np
import pytest

from flypylib_b200 import fplobjdetect as F
from flypylib_b200 import fplutils
from oracle import voxel2obj_oracle as O
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "substack_golden.npz"))


@pytest.mark.parametrize("name", ["inside", "corner", "far", "outside"])
def test_fri_get_image_matches_reference(name, tmp_path):
    size, z, y, x, buf = (int(v) for v in GOLD[name + "/spec"])
    ss = fplutils.szyx(size, z, y, x)
    img, back = F.fri_get_image([ss, list(GOLD[name + "/norm"]), buf, str(tmp_path)], GOLD["volume"])
    assert back == ss
    if bool(GOLD[name + "/none"]):
        assert img is None
        return
    want = GOLD[name + "/image"]
    assert img.dtype == want.dtype and np.array_equal(img, want)
    txt = open("%s/%d_%d_%d_%d.txt" % (tmp_path, size, z, y, x)).read()
    assert txt == GOLD[name + "/txt"].tobytes().decode()


class FakeNet(object):
    """network.infer stand-in: a smooth positive function of the normalised image (no GPU)."""

    def __init__(self):
        self.calls = 0

    def infer(self, image):
        self.calls += 1
        a = np.asarray(image, dtype=np.float32)
        return (1.0 / (1.0 + np.exp(-a))).astype(np.float32)


def _oracle_v2o(pred, r, sigma, off=(0, 0, 0), buf=0, thd=0, **kw):
    return O.voxel2obj(np.asarray(pred, dtype=np.float32), r, sigma, off, buf, thd, impl="c")


def _roi_file(path, substacks):
    with open(path, "w") as f:
        f.write("\n".join("%d,%d,%d,%d" % s for s in substacks) + "\n")


SUBSTACKS = [(32, 0, 0, 0), (32, 0, 0, 32), (32, 32, 0, 0), (32, 32, 32, 32), (32, 200, 0, 0)]


def _expected(vol, substacks, buf, norm, r, sigma, thd, tmp):
    locs, conf = [], []
    net = FakeNet()
    for s in substacks:
        ss = fplutils.szyx(*s)
        img, _ = F.fri_get_image([ss, norm, buf, str(tmp)], vol)
        if img is None:
            continue
        o = _oracle_v2o(net.infer(img), r, sigma, (ss.x - buf, ss.y - buf, ss.z - buf), buf, thd)
        locs.append(o["locs"]); conf.append(o["conf"])
    return np.concatenate(locs), np.concatenate(conf)


def test_full_roi_inference_flow_and_resume(tmp_path, monkeypatch):
    monkeypatch.setattr(F, "voxel2obj", _oracle_v2o)
    vol = cases.em_volume((72, 70, 76), seed=3)
    roi = str(tmp_path / "roi.txt"); _roi_file(roi, SUBSTACKS)
    wd = str(tmp_path / "work")
    norm = [128., 33., 0.5]
    net = FakeNet()
    out = F.full_roi_inference(vol, None, roi, net, 0, wd, norm, obj_min_dist=5, smoothing_sigma=1.5, buffer_sz=6)
    os.makedirs(str(tmp_path / "exp"), exist_ok=True)
    el, ec = _expected(vol, SUBSTACKS, 6, norm, 5, 1.5, 0, tmp_path / "exp")
    assert out["conf"].size > 10 and net.calls == 4          # the substack outside the volume is skipped
    assert np.array_equal(out["locs"], el) and np.array_equal(out["conf"], ec)
    # detections lie inside their substack (buffer zone dropped) and carry the global offset
    for s in SUBSTACKS:
        assert os.path.isfile(F.fri_filename(wd, fplutils.szyx(*s)))
    with open(wd + "/all.p", "rb") as f:
        allp = pickle.load(f)
    assert np.array_equal(allp["locs"], out["locs"])
    # resume: nothing is recomputed, same result
    net2 = FakeNet()
    out2 = F.full_roi_inference(vol, None, roi, net2, 0, wd, norm, obj_min_dist=5, smoothing_sigma=1.5, buffer_sz=6)
    assert net2.calls == 0
    assert np.array_equal(out2["locs"], out["locs"]) and np.array_equal(out2["conf"], out["conf"])
    # partial resume after losing one pickle
    os.unlink(F.fri_filename(wd, fplutils.szyx(*SUBSTACKS[1])))
    net3 = FakeNet()
    out3 = F.full_roi_inference(vol, None, roi, net3, 0, wd, norm, obj_min_dist=5, smoothing_sigma=1.5, buffer_sz=6)
    assert net3.calls == 1 and np.array_equal(out3["locs"], out["locs"])


def test_full_roi_inference_rejects_out_of_scope_sources(tmp_path):
    roi = str(tmp_path / "roi.txt"); _roi_file(roi, SUBSTACKS[:1])
    with pytest.raises(NotImplementedError):
        F.full_roi_inference("http://dvid:8000", "abc", roi, FakeNet(), 0, str(tmp_path / "w"), [128., 33.])
    with pytest.raises(NotImplementedError):
        F.full_roi_inference(np.zeros((8, 8, 8), np.uint8), None, roi, FakeNet(), 0, str(tmp_path / "w"), [128., 33.],
                             dvid_seg_info=("s", "u"))


def _worker(rank, world, port, tmp, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F.voxel2obj = _oracle_v2o
    vol = cases.em_volume((72, 70, 76), seed=3)
    net = FakeNet()
    out = F.full_roi_inference(vol, None, tmp + "/roi.txt", net, 0, tmp + "/work2", [128., 33., 0.5],
                               obj_min_dist=5, smoothing_sigma=1.5, buffer_sz=6)
    q.put((rank, net.calls, out["locs"], out["conf"]))
    dist.barrier()
    dist.destroy_process_group()


def test_full_roi_inference_world2_gloo(tmp_path):
    """Two ranks (gloo): the pending substacks are dealt round-robin, both ranks return the complete list."""
    import torch.multiprocessing as mp
    _roi_file(str(tmp_path / "roi.txt"), SUBSTACKS)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=180) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
    os.makedirs(str(tmp_path / "exp"), exist_ok=True)
    el, ec = _expected(cases.em_volume((72, 70, 76), seed=3), SUBSTACKS, 6, [128., 33., 0.5], 5, 1.5, 0, tmp_path / "exp")
    assert res[0][1] + res[1][1] == 4 and res[0][1] >= 1 and res[1][1] >= 1
    for r in res:
        assert np.array_equal(r[2], el) and np.array_equal(r[3], ec)


@pytest.mark.gpu
def test_full_roi_inference_gpu_end_to_end(tmp_path):
    """Real network + CUDA voxel2obj through the driver == the same substacks processed by hand."""
    import torch
    import bench
    from flypylib_b200 import fplmodels, fplnetwork
    net = fplnetwork.FplNetwork(fplmodels.vgg_like2)
    net.train_single.set_weights(bench.seeded_weights("vgg_like2"))
    net.set_precision("bf16")
    net._set_infer()
    vol = cases.em_volume((150, 140, 160), seed=4)
    subs = [(64, 0, 0, 0), (64, 64, 64, 64), (64, 64, 0, 96)]
    roi = str(tmp_path / "roi.txt"); _roi_file(roi, subs)
    out = F.full_roi_inference(vol, None, roi, net, 0, str(tmp_path / "w"), [128., 33.], buffer_sz=30)
    locs, conf = [], []
    for s in subs:
        ss = fplutils.szyx(*s)
        img, _ = F.fri_get_image([ss, [128., 33.], 30, str(tmp_path)], vol)
        pred = net.infer_device(torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).cuda())
        o = F.voxel2obj(pred, 27, 5, (ss.x - 30, ss.y - 30, ss.z - 30), 30, 0)
        locs.append(o["locs"]); conf.append(o["conf"])
    assert out["conf"].size > 3
    assert np.array_equal(out["locs"], np.concatenate(locs)) and np.array_equal(out["conf"], np.concatenate(conf))
    lo = np.array([0, 0, 0]); hi = np.array([160, 140, 150])
    assert np.all(out["locs"] >= lo) and np.all(out["locs"] < hi)


def test_evaluate_substacks_inline(tmp_path, monkeypatch):
    """evaluate_substacks (fplobjdetect.py:484-512): per-substack PR curves and their aggregate; checked against
    the pieces called by hand (oracle voxel2obj standing in for the CUDA one, fake network)."""
    from flypylib_b200 import fplsynapses
    monkeypatch.setattr(F, "voxel2obj", _oracle_v2o)
    net = FakeNet()
    thds = np.array([0.0, 0.5, 0.6, 0.9])
    subs, expected = [], []
    for seed in (1, 2):
        img = ((cases.em_volume((40, 44, 48), seed=seed).astype(np.float32) - 128.0) / 33.0)
        pred = net.infer(img)
        dets = _oracle_v2o(pred, 5, 1.5, (0, 0, 0), 4, 0)
        keep = np.arange(dets["conf"].size) % 3 != 0               # ground truth = two thirds of the detections
        gt = {"locs": dets["locs"][keep] + 1.0, "conf": dets["conf"][keep]}
        jf = str(tmp_path / ("gt%d.json" % seed))
        fplsynapses.tbars_to_json_format(gt, json_file=jf)
        subs.append((img, jf))
        expected.append(F.obj_pr_curve(dets, fplsynapses.load_from_json(jf, pred.shape, 4), 5, thds))
    agg, results = F.evaluate_substacks(net, subs, thds, obj_min_dist=5, smoothing_sigma=1.5, buffer_sz=4)
    assert len(results) == 2
    for r, e in zip(results, expected):
        for k in ("num_tp", "tot_pred", "tot_gt", "pp", "rr"):
            assert np.array_equal(getattr(r, k), getattr(e, k)), k
    want = F.aggregate_pr(expected)
    assert np.array_equal(agg.num_tp, want.num_tp) and np.array_equal(agg.pp, want.pp)
    assert agg.num_tp[0] > 0
