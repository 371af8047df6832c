"""CPU: detection scoring (obj_match / obj_pr / obj_pr_curve / aggregate_pr) and the json wire
formats against goldens produced by the UNMODIFIED reference functions
(tests/golden/make_golden.py:golden_eval; flypylib/fplobjdetect.py:259-453, fplsynapses.py:11-111)."""
import json
import os

import numpy as np
import pytest

from flypylib_b200 import fplobjdetect as F
from flypylib_b200 import fplsynapses as S
from oracle import match_oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eval_golden.json")))


@pytest.mark.parametrize("idx", range(len(GOLD["pr"])))
def test_obj_pr_and_curve_match_reference(idx):
    g = GOLD["pr"][idx]
    pred, conf, gt = np.array(g["pred"]), np.array(g["conf"]), np.array(g["gt"])
    r = F.obj_pr(pred, gt, 27.0, allow_mult=g["allow_mult"])
    assert (int(r.num_tp), int(r.tot_pred), int(r.tot_gt)) == (g["num_tp"], g["tot_pred"], g["tot_gt"])
    assert r.pp == g["pp"] and r.rr == g["rr"]
    d = np.sqrt(((pred[:, None] - gt[None]) ** 2).sum(2)) - 27.0
    assert abs(d[r.match].sum() - g["match_cost"]) < 1e-9
    c = F.obj_pr_curve({"locs": pred, "conf": conf}, {"locs": gt}, 27.0, np.array(GOLD["thresholds"]),
                       allow_mult=g["allow_mult"])
    for k, v in g["curve"].items():
        assert np.array_equal(np.asarray(getattr(c, k)), np.asarray(v)), k


def test_aggregate_pr_matches_reference():
    th = np.array(GOLD["thresholds"])
    curves = [F.obj_pr_curve({"locs": np.array(g["pred"]), "conf": np.array(g["conf"])}, {"locs": np.array(g["gt"])},
                             27.0, th) for g in GOLD["pr"] if not g["allow_mult"]]
    agg = F.aggregate_pr(curves)
    for k, v in GOLD["aggregate"].items():
        assert np.array_equal(np.asarray(getattr(agg, k)), np.asarray(v)), k


def test_empty_sides():
    e = F.obj_pr(np.zeros((0, 3)), np.zeros((4, 3)), 27.0)
    assert [int(e.num_tp), int(e.tot_pred), int(e.tot_gt), e.pp, e.rr, e.match] == GOLD["empty_pred"]
    e = F.obj_pr(np.zeros((3, 3)), np.zeros((0, 3)), 27.0)
    assert [int(e.num_tp), int(e.tot_pred), int(e.tot_gt), e.pp, e.rr, e.match] == GOLD["empty_gt"]


@pytest.mark.parametrize("allow_mult", [False, True])
def test_obj_match_optimal_vs_exhaustive_oracle(allow_mult):
    rng = np.random.default_rng(11)
    for _ in range(150):
        n, m = rng.integers(1, 9, 2)
        d = rng.normal(0.3, 1.0, (n, m))
        a, b = match_oracle.obj_match(d, allow_mult), F.obj_match(d, allow_mult)
        assert abs(d[a].sum() - d[b].sum()) < 1e-12 and a.sum() == b.sum()
        assert (d[b] < 0).all() and (b.sum(0) <= 1).all()
        if not allow_mult:
            assert (b.sum(1) <= 1).all()


def test_obj_match_large_sparse_is_fast_and_valid():
    rng = np.random.default_rng(5)
    gt = rng.uniform(0, 2000, (4000, 3))
    pred = np.concatenate([gt[:3500] + rng.normal(0, 5, (3500, 3)), rng.uniform(0, 2000, (600, 3))])
    r = F.obj_pr(pred, gt, 27.0)
    assert 3300 <= r.num_tp <= 3600 and (r.match.sum(0) <= 1).all() and (r.match.sum(1) <= 1).all()


def test_label_constraint_blocks_cross_label_matches():
    pred = np.array([[0., 0, 0], [10, 0, 0]]); gt = np.array([[1., 0, 0], [11, 0, 0]])
    r = F.obj_pr(pred, gt, 5.0, predict_lbls=np.array([1, 2]), groundtruth_lbls=np.array([1, 3]))
    assert r.num_tp == 1 and r.match[0, 0] and not r.match[1, 1]


def _unpack(t):
    return {k: np.asarray(v) for k, v in t.items()}


def test_json_writers_match_reference(tmp_path):
    f = GOLD["formats"]
    tb = {"locs": np.array(f["locs"]), "conf": np.array(f["conf"])}
    assert S.tbars_to_json_format(tb, labels=np.arange(9) * 7) == f["dvid_labels"]
    assert S.tbars_to_json_format(tb, user_name="someone") == f["dvid_user"]
    assert S.tbars_to_json_format_raveler(tb) == f["raveler"]
    p = str(tmp_path / "t.json")
    S.tbars_to_json_format(tb, json_file=p)
    assert open(p).read() == f["file_text"]
    back = S.load_from_json(p)
    assert np.array_equal(back["locs"], tb["locs"].astype(int))


@pytest.mark.parametrize("key,src,kw", [("back_dvid", "mixed", {}), ("back_nested", "nested", {}),
                                        ("back_raveler", "raveler", {}),
                                        ("back_buffer", "dvid_user", dict(vol_sz=500, buffer=(60, 40, 80)))])
def test_load_from_json_matches_reference(key, src, kw):
    f = GOLD["formats"]
    doc = [f["dvid_user"]] if src == "nested" else f[src]
    got = S.load_from_json(json.dumps(doc), **kw)
    want = _unpack(f[key])
    assert set(got) == set(want)
    for k in want:
        g, w = np.asarray(got[k]), want[k]
        assert g.shape == w.shape, k
        assert [None if x is None else x for x in g.tolist()] == w.tolist(), k


@pytest.mark.parametrize("tag,ctx,bs,is_mask", [("cls", (8, 8, 8), 6, False), ("mask", (10, 10, 10), 4, True)])
def test_gen_batches_matches_reference(tag, ctx, bs, is_mask):
    """gen_batches (fplobjdetect.py:27-130): with the same np.random seed the batches equal those of the
    unmodified reference bit for bit (tests/golden/make_golden.py:golden_gen_batches)."""
    from tests.golden import make_golden
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "gen_batches_golden.npz"))
    np.random.seed(1234)
    g = F.gen_batches(make_golden.gen_batches_volumes(), ctx, bs, is_mask=is_mask)
    for k in range(3):
        d, l = next(g)
        assert d.dtype == np.float32 and l.dtype == np.uint8
        assert np.array_equal(d, gold["%s/data%d" % (tag, k)]), (tag, k)
        assert np.array_equal(l, gold["%s/labels%d" % (tag, k)]), (tag, k)
