"""Drop-in for the detection half of flypylib/fplobjdetect.py: ``voxel2obj``.

Same signature, same ``{'locs','conf'}`` result (float64, (x,y,z) columns, emission order) as
flypylib/fplobjdetect.py:132-257 -- but every array operation runs as sm_100a CUDA kernels behind
the C ABI (include/fpl_b200.h, fpl_voxel2obj).  The host code below only derives the scalars the
reference derives on the host: SciPy's Gaussian taps, NumPy's float32 percentile index arithmetic
and the promotion of ``thd``.
"""
import ctypes

import numpy as np

from . import _lib
from . import fplutils


def _gaussian_taps(sigma, truncate=2.0):
    """Taps scipy.ndimage.gaussian_filter(.., sigma, truncate=2.0) uses
    (scipy _filters.py gaussian_filter1d/_gaussian_kernel1d; reference call fplobjdetect.py:167-168)."""
    sd = float(sigma)
    if not sd > 1e-15:            # SciPy skips axes with sigma <= 1e-15
        return None, -1
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[::-1], dtype=np.float64), lw


def _percentile_plan(n, q=97):
    """(rank_lo, rank_hi, gamma) of np.percentile(<float32 array of n values>, q), method 'linear'.

    NumPy (>= 2.0) keeps the whole index computation in float32 for float32 data:
    q/100 -> float32, virtual index (n-1)*q -> float32, floor, +1, gamma = vi - floor(vi).
    The reference calls np.percentile(pred, 97) on the padded float32 map (fplobjdetect.py:183).
    """
    qf = np.true_divide(q, np.float32(100))
    vi = np.asanyarray((n - 1) * qf)
    prev = np.asanyarray(np.floor(vi))
    nxt = np.asanyarray(prev + 1)
    if vi >= n - 1:
        prev = nxt = np.asanyarray(np.float32(-1))
    if vi < 0:
        prev = nxt = np.asanyarray(np.float32(0))
    gamma = np.asanyarray(vi - prev, dtype=vi.dtype)
    lo = int(prev.astype(np.intp))
    hi = int(nxt.astype(np.intp))
    if lo < 0:
        lo += n
    if hi < 0:
        hi += n
    return lo, hi, float(gamma)


def _promote_thd(thd):
    """Value of ``thd`` after np.maximum(<float32 scalar>, thd) style promotion, as a python float."""
    t = np.maximum(np.float32(-np.inf), thd)
    return float(t)


def _make_params(shape, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd):
    r = int(obj_min_dist)
    if r != obj_min_dist:
        raise TypeError("obj_min_dist must be an integer (the reference uses it as a pad width)")
    Z, Y, X = (int(s) for s in shape)
    buf = fplutils.to3d(buffer_sz)
    weights, lw = _gaussian_taps(smoothing_sigma)
    n_pad = (Z + 2 * r) * (Y + 2 * r) * (X + 2 * r)
    lo, hi, gamma = _percentile_plan(n_pad, 97)
    p = _lib.V2OParams()
    p.obj_min_dist = r
    p.lw = lw
    p.h_weights = weights.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if weights is not None else None
    p.thd = _promote_thd(thd)
    p.rank_lo, p.rank_hi, p.gamma = lo, hi, gamma
    for i in range(3):
        b = buf[i]
        if int(b) != b:
            raise TypeError("buffer_sz must be integral")
        p.buffer_xyz[i] = int(b)
    off = tuple(volume_offset)          # the reference requires a tuple (fplobjdetect.py:252)
    if len(off) != 3:
        raise ValueError("volume_offset must have 3 entries (x,y,z)")
    for i in range(3):
        p.offset_xyz[i] = float(off[i])
    return p, weights


def _default_capacity(shape, r):
    """Upper bound on detections: points are pairwise > r apart, so balls of radius r/2 are
    disjoint; use a generous closed-form bound with slack."""
    Z, Y, X = shape
    rr = max(r, 1)
    cell = max(1.0, rr / 2.0)
    bound = (Z / cell + 2) * (Y / cell + 2) * (X / cell + 2)
    return int(min(Z * Y * X, bound)) + 64


def voxel2obj_device(pred_dev, obj_min_dist, smoothing_sigma, volume_offset=(0, 0, 0), buffer_sz=0,
                     thd=0, capacity=None, return_stats=False):
    """voxel2obj on a CUDA float32 tensor (Z,Y,X); returns the dict plus optional stats."""
    import torch
    if not (isinstance(pred_dev, torch.Tensor) and pred_dev.is_cuda):
        raise TypeError("voxel2obj_device expects a CUDA tensor")
    if pred_dev.dtype != torch.float32 or pred_dev.dim() != 3:
        raise TypeError("voxel2obj_device expects a 3-D float32 tensor")
    pred_dev = pred_dev.contiguous()
    dev = pred_dev.device.index
    ctx = _lib.context(dev)
    shape = tuple(pred_dev.shape)
    p, _keepalive = _make_params(shape, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd)
    cap = int(capacity) if capacity is not None else _default_capacity(shape, p.obj_min_dist)
    with torch.cuda.device(dev):
        dets = torch.empty((cap, 4), dtype=torch.float64, device=pred_dev.device)
        count = ctypes.c_int64(0)
        thresh = ctypes.c_double(0)
        stats = (ctypes.c_int64 * 8)()
        rc = _lib.lib().fpl_voxel2obj(ctx.handle, pred_dev.data_ptr(), shape[0], shape[1], shape[2],
                                     ctypes.byref(p), dets.data_ptr(), cap, ctypes.byref(count),
                                     ctypes.byref(thresh), stats, _lib.current_stream_ptr(dev))
        _lib.check(rc, "fpl_voxel2obj")
        rows = dets[:count.value].cpu().numpy()
    out = {'locs': rows[:, :3].copy(), 'conf': rows[:, 3].copy()}
    if return_stats:
        # stats[6]: which path of fpl_voxel2obj produced the result (all three are bit-identical by construction)
        path = {2: 'two-tier', 1: 'fused-exact'}.get(int(stats[6]), 'classic-exact')
        lib = _lib.lib()
        lib.fpl_debug_v2o_decline_reason.argtypes = [ctypes.c_void_p, ctypes.c_int]
        declined = int(lib.fpl_debug_v2o_decline_reason(ctx.handle, 0))
        return out, {'threshold': thresh.value, 'candidates': stats[0], 'rounds': stats[1],
                     'ball_checks': stats[2], 'selected': stats[3], 'path': path, 'two_tier_declined': declined,
                     'exact_recomputed': int(stats[4]) if path == 'two-tier' else None,
                     'ambiguous_ball_checks': int(stats[7]) if path == 'two-tier' else None}
    return out


def voxel2obj(pred, obj_min_dist, smoothing_sigma,
              volume_offset=(0, 0, 0), buffer_sz=0, thd=0,
              seg=None, seg_dilate=None, seg_sz_thd=None,
              seg_force=None):
    """convert voxel-wise predictions to object predictions (flypylib/fplobjdetect.py:132-257).

    Args / returns exactly as the reference: ``pred`` 3-D array of voxel-wise predictions (numpy,
    or a CUDA torch tensor to skip the host->device copy), ``obj_min_dist`` suppression radius,
    ``smoothing_sigma`` Gaussian sigma, ``volume_offset`` (x,y,z) shift, ``buffer_sz`` border in
    which detections are dropped, ``thd`` lower bound on the threshold.  Returns
    ``{'locs': (N,3) float64 (x,y,z), 'conf': (N,) float64}`` in the reference's emission order.

    With ``seg`` (label volume of ``pred``'s shape, array or h5 path) the segmentation-aware branch runs
    (:161-165,177-181,192-195,213-224): ``seg_sz_thd`` zeroes the smoothed map inside segments smaller than that many
    voxels, a selected point suppresses only ball voxels inside its own segment dilated ``seg_dilate`` times
    (6-connected), ``seg_force`` always suppresses an inner ball -- see ``voxel2obj_seg_device``.
    """
    import torch
    if seg is not None or seg_dilate is not None or seg_sz_thd is not None or seg_force:
        return _voxel2obj_seg(pred, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd, seg, seg_dilate,
                              seg_sz_thd, seg_force)
    if isinstance(pred, str):                            # fplobjdetect.py:154-156: h5 file with the map in /main
        from . import h5lite
        pred = h5lite.File(pred)['/main'][:]
    if isinstance(pred, torch.Tensor):
        dev = pred if pred.is_cuda else pred.cuda()
        if dev.dtype != torch.float32:
            raise TypeError("voxel2obj on device supports float32 probability maps only")
    else:
        a = np.asarray(pred)
        if a.ndim != 3:
            raise ValueError("pred must be 3-D")
        if a.dtype != np.float32:
            raise TypeError("the B200 voxel2obj path is the float32 path FplNetwork.infer produces; "
                            "got dtype %s" % a.dtype)
        _lib.context()          # raises when there is no GPU: no CPU fallback
        dev = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return voxel2obj_device(dev, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd)


def _voxel2obj_seg(pred, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd, seg, seg_dilate, seg_sz_thd,
                   seg_force):
    import torch
    from . import h5lite
    if isinstance(pred, str):
        pred = h5lite.File(pred)['/main'][:]
    if isinstance(seg, str):                             # fplobjdetect.py:162-164
        seg = h5lite.File(seg)['/main'][:]
    _lib.context()                                       # raises when there is no GPU: no CPU fallback
    if isinstance(pred, torch.Tensor):
        pdev = pred if pred.is_cuda else pred.cuda()
    else:
        a = np.asarray(pred)
        if a.ndim != 3:
            raise ValueError("pred must be 3-D")
        if a.dtype != np.float32:
            raise TypeError("the B200 voxel2obj path is the float32 path FplNetwork.infer produces; got dtype %s" % a.dtype)
        pdev = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sdev = None
    if seg is not None:
        if isinstance(seg, torch.Tensor):
            sdev = seg.to(device=pdev.device, dtype=torch.int64)
        else:
            s = np.asarray(seg)
            if s.dtype == np.uint64:
                s = s.view(np.int64)                     # labels are only compared for equality
            sdev = torch.from_numpy(np.ascontiguousarray(s.astype(np.int64, copy=False))).to(pdev.device)
        if tuple(sdev.shape) != tuple(pdev.shape):
            raise ValueError("seg must have the shape of pred")
    elif seg_sz_thd is not None:
        raise TypeError("seg_sz_thd needs a segmentation")           # np.unique(None) in the reference
    return voxel2obj_seg_device(pdev, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd, sdev, seg_dilate,
                                seg_sz_thd, seg_force)


def voxel2obj_seg_device(pred_dev, obj_min_dist, smoothing_sigma, volume_offset=(0, 0, 0), buffer_sz=0, thd=0,
                         seg_dev=None, seg_dilate=None, seg_sz_thd=None, seg_force=None):
    """Segmentation-aware voxel2obj on device tensors (pred float32 (Z,Y,X), seg int64 (Z,Y,X) or None).

    Stages: exact smoothing (``fpl_v2o_smooth``) -> optional size filter (:177-181: label counts of the zero-PADDED
    segmentation; device-side ``torch.unique``) -> threshold (``fpl_v2o_threshold``) -> candidates sorted by (value desc,
    index asc) -> ``fpl_v2o_detect_seg``: the reference's sequential loop with the segment-shaped suppression masks,
    in one persistent CTA.  Everything stays on the GPU; bit-exact against the unmodified reference
    (tests/golden/voxel2obj_seg_golden.npz)."""
    import torch
    if seg_dilate is not None and int(seg_dilate) < 1:
        raise ValueError("seg_dilate must be >= 1 (SciPy's iterations < 1 means 'until convergence')")
    pred_dev = pred_dev.contiguous()
    dev = pred_dev.device.index
    ctx = _lib.context(dev)
    lib = _lib.lib()
    Z, Y, X = (int(v) for v in pred_dev.shape)
    p, _keepalive = _make_params((Z, Y, X), obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd)
    r = int(p.obj_min_dist)
    with torch.cuda.device(dev):
        st = _lib.current_stream_ptr(dev)
        smooth = torch.empty_like(pred_dev)
        _lib.check(lib.fpl_v2o_smooth(ctx.handle, pred_dev.data_ptr(), Z, Y, X, ctypes.byref(p), smooth.data_ptr(), st),
                   "fpl_v2o_smooth")
        if seg_sz_thd is not None:
            ids, inv, counts = torch.unique(seg_dev, return_inverse=True, return_counts=True)
            counts = counts.clone()
            pad_vox = (Z + 2 * r) * (Y + 2 * r) * (X + 2 * r) - Z * Y * X
            counts[ids == 0] += pad_vox                  # the zero padding belongs to label 0
            small = counts < int(seg_sz_thd)
            smooth[small[inv]] = 0.0
            del ids, inv, counts, small
        h = (ctypes.c_double * 4)()
        _lib.check(lib.fpl_v2o_threshold(ctx.handle, smooth.data_ptr(), Z, Y, X, ctypes.byref(p), h, st),
                   "fpl_v2o_threshold")
        thr = float(h[0])
        # float32 v > thr (float64)  <=>  v > (largest float32 <= thr)
        t32 = np.float32(thr)
        if float(t32) > thr:
            t32 = np.nextafter(t32, np.float32(-np.inf))
        flat = smooth.reshape(-1)
        idx = torch.nonzero(flat > float(t32)).reshape(-1)
        vals, order = torch.sort(flat[idx], descending=True, stable=True)     # ties keep ascending index order
        idx = idx[order].contiguous()
        n = Z * Y * X
        supp = torch.zeros(((n + 31) // 32,), dtype=torch.int32, device=pred_dev.device)
        # with a segmentation two detections may be neighbours (different segments): every candidate can be selected
        cap = max(int(idx.numel()), 1) if seg_dev is not None else min(_default_capacity((Z, Y, X), r), max(int(idx.numel()), 1))
        rows = torch.empty((cap, 4), dtype=torch.float64, device=pred_dev.device)
        count = torch.zeros((3,), dtype=torch.int64, device=pred_dev.device)
        _lib.check(lib.fpl_v2o_detect_seg(ctx.handle, vals.data_ptr(), idx.data_ptr(), int(idx.numel()),
                                          seg_dev.data_ptr() if seg_dev is not None else None, supp.data_ptr(), Z, Y, X,
                                          ctypes.byref(p), -1 if seg_dilate is None else int(seg_dilate),
                                          int(seg_force) if seg_force else 0, rows.data_ptr(), cap, count.data_ptr(), st),
                   "fpl_v2o_detect_seg")
        c = count.cpu().numpy()
        if c[2]:
            raise _lib.FplError("voxel2obj (seg): detection list overflow (%d rows)" % cap)
        out = rows[:int(c[0])].cpu().numpy()
    return {'locs': out[:, :3].copy(), 'conf': out[:, 3].copy()}


# ---------------------------------------------------------------------------------------------
# Detection scoring: obj_match / obj_pr / obj_pr_curve / aggregate_pr.
# Host-side evaluation, not part of the GPU hot path; it exists so that the "bf16 keeps detection
# F1 within 0.5 %" criterion can be measured with the reference's own definition of a match.
# The reference solves the matching as an integer program through PuLP (fplobjdetect.py:259-318);
# PuLP is absent here, and the program is a bipartite minimum-cost partial matching, solved exactly
# per connected component of the admissible-pair graph with the Hungarian method.
# ---------------------------------------------------------------------------------------------
import collections

PR_Result = collections.namedtuple('PR_Result', 'num_tp tot_pred tot_gt pp rr match')


def obj_match(dists, allow_mult=False):
    """Optimal prediction/ground-truth matching (flypylib/fplobjdetect.py:259-318).

    ``dists`` is the (N,M) distance matrix minus the distance threshold: negative entries are the
    admissible pairs.  Minimises the summed entry over the chosen pairs subject to "each ground
    truth at most once" and, unless ``allow_mult``, "each prediction at most once".  Returns the
    (N,M) boolean match matrix.  Every admissible cost is negative, so with ``allow_mult`` the
    optimum is simply: each ground-truth column takes its best admissible prediction.
    """
    from scipy.optimize import linear_sum_assignment
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    d = np.asarray(dists, dtype=np.float64)
    n_pred, n_gt = d.shape
    out = np.zeros((n_pred, n_gt), dtype=bool)
    ok = d < 0
    if not ok.any():
        return out
    if allow_mult:
        masked = np.where(ok, d, np.inf)
        best = masked.argmin(axis=0)
        cols = np.nonzero(ok.any(axis=0))[0]
        out[best[cols], cols] = True
        return out
    ii, jj = np.nonzero(ok)
    graph = coo_matrix((np.ones(ii.size, np.int8), (ii, jj + n_pred)), shape=(n_pred + n_gt,) * 2)
    _, comp = connected_components(graph, directed=False)
    order = np.argsort(comp[ii], kind='stable')
    ii, jj = ii[order], jj[order]
    bounds = np.flatnonzero(np.diff(comp[ii])) + 1
    for ri, ci in zip(np.split(ii, bounds), np.split(jj, bounds)):
        rows, cols = np.unique(ri), np.unique(ci)
        if rows.size == 1 or cols.size == 1:                 # star: the single best pair wins
            k = d[ri, ci].argmin()
            out[ri[k], ci[k]] = True
            continue
        sub = d[np.ix_(rows, cols)]
        cost = np.where(sub < 0, sub, 0.0)                   # 0 == "leave unmatched"
        a, b = linear_sum_assignment(cost)
        keep = cost[a, b] < 0
        out[rows[a[keep]], cols[b[keep]]] = True
    return out


def obj_pr(predict_locs, groundtruth_locs, dist_thresh, predict_lbls=None, groundtruth_lbls=None,
           allow_mult=False):
    """Precision / recall of predicted against ground-truth locations (fplobjdetect.py:320-374).

    Returns ``PR_Result(num_tp, tot_pred, tot_gt, pp, rr, match)`` with the reference's conventions:
    an empty side gives ``pp = 1`` when there are no predictions and ``rr = 1`` when there is no
    ground truth, ``match = None``; otherwise ``pp = num_tp / N`` and ``rr = num_tp / M`` while
    ``tot_pred`` additionally counts the surplus matches of multiply-matched predictions.
    """
    predict_locs = np.asarray(predict_locs)
    groundtruth_locs = np.asarray(groundtruth_locs)
    n_pred, n_gt = predict_locs.shape[0], groundtruth_locs.shape[0]
    if n_pred == 0 or n_gt == 0:
        return PR_Result(num_tp=0, tot_pred=n_pred, tot_gt=n_gt, pp=1 if n_pred == 0 else 0,
                         rr=1 if n_gt == 0 else 0, match=None)
    diff = predict_locs.reshape((-1, 1, 3)) - groundtruth_locs.reshape((1, -1, 3))
    dists = np.sqrt((diff ** 2).sum(axis=2))
    dists -= dist_thresh
    if predict_lbls is not None:
        differ = (np.reshape(predict_lbls, (-1, 1)) != np.reshape(groundtruth_lbls, (1, -1)))
        dists += (dist_thresh + 1.) * differ.astype('float32')
    match = obj_match(dists, allow_mult=allow_mult)
    num_tp = match.sum()
    surplus = np.maximum(match.sum(axis=1) - 1, 0).sum()
    return PR_Result(num_tp=num_tp, tot_pred=n_pred + surplus, tot_gt=n_gt, pp=num_tp / n_pred,
                     rr=num_tp / n_gt, match=match)


def obj_pr_curve(predict, groundtruth, dist_thresh, thresholds, predict_lbls=None,
                 groundtruth_lbls=None, allow_mult=False):
    """Precision / recall at each confidence threshold (fplobjdetect.py:376-434): predictions with
    ``conf >= thresholds[t]`` are scored by :func:`obj_pr`; ``match`` is the first threshold's matrix.
    ``predict`` / ``groundtruth`` are ``{'locs','conf'}`` dicts or paths of json files written by
    ``fplsynapses.tbars_to_json_format``."""
    from . import fplsynapses
    if isinstance(predict, str):
        predict = fplsynapses.load_from_json(predict)
    if isinstance(groundtruth, str):
        groundtruth = fplsynapses.load_from_json(groundtruth)
    thresholds = np.asarray(thresholds)
    cols = {k: np.zeros((thresholds.size,)) for k in ('num_tp', 'tot_pred', 'tot_gt', 'pp', 'rr')}
    match = None
    for t in range(thresholds.size):
        sel = predict['conf'] >= thresholds[t]
        lbls = predict_lbls[sel] if predict_lbls is not None else None
        res = obj_pr(predict['locs'][sel, :], groundtruth['locs'], dist_thresh, lbls, groundtruth_lbls,
                     allow_mult=allow_mult)
        for k in cols:
            cols[k][t] = getattr(res, k)
        if match is None:
            match = res.match
    return PR_Result(match=match, **cols)


def aggregate_pr(results):
    """Pool per-substack PR curves (fplobjdetect.py:437-453): counts add, precision and recall are
    recomputed from the pooled counts with the reference's 10e-8 guard."""
    num_tp = np.zeros(results[0].num_tp.shape)
    tot_pred = np.zeros_like(num_tp)
    tot_gt = np.zeros_like(num_tp)
    for res in results:
        num_tp += res.num_tp
        tot_pred += res.tot_pred
        tot_gt += res.tot_gt
    return PR_Result(num_tp=num_tp, tot_pred=tot_pred, tot_gt=tot_gt, pp=num_tp / (tot_pred + 10e-8),
                     rr=num_tp / (tot_gt + 10e-8), match=None)


# ---------------------------------------------------------------------------------------------
# Substack driver: full_roi_inference and its fri_* helpers
# (flypylib/fplobjdetect.py:841-1059) -- the production caller on both sides of the hot path.
# Kept: ROI text file -> szyx substacks, size + 2*buffer cubes zero-padded at the volume faces,
# the uint8 normalisation with `global_frac`, the per-substack norm .txt and result pickles (resume),
# voxel2obj per substack with the substack offset and buffer, all.p.  Changed, B200-first: no worker
# processes (voxel2obj cannot run in forked children of a CUDA process -- it runs inline, on the
# probability map that is still resident in HBM); with torch.distributed initialised the substacks are
# dealt round-robin to the ranks (one process per GPU) and the detection lists are all-gathered.
# Out of scope: DVID / DICED / N5 volume services (libdvid, diced, z5py are storage/network control
# plane) -- `data_source` is any (Z,Y,X) uint8 array-like that supports slicing (numpy array,
# numpy.memmap, or an object with .shape and __getitem__), which is what the reference's
# "n5://" branch reduces the store to (fplobjdetect.py:1040-1068).
# ---------------------------------------------------------------------------------------------
import os
import pickle
import sys

from .fplutils import szyx, roi_from_txt  # noqa: E402,F401


def fri_filename(working_dir, substack):
    """Result pickle of a substack (fplobjdetect.py:1153-1155)."""
    return '%s/%d_%d_%d_%d.p' % (working_dir, substack.size, substack.z, substack.y, substack.x)


def fri_get_image(substack_info, volume):
    """Cut the (size + 2*buffer)^3 cube of a substack out of `volume` (zeros beyond its faces) and
    normalise it as the reference does (fplobjdetect.py:1021-1112): mean of the voxels in (1, 200)
    blended with the global mean by ``global_frac`` = image_normalize[2] (1 when absent), divided by
    image_normalize[1]; the statistics go to ``<norm_dir>/<size>_<z>_<y>_<x>.txt``.
    Returns ``(float32 image or None, substack)``; None when the cube lies outside the volume."""
    substack, image_normalize, buffer_sz, norm_dir = substack_info
    image_sz = substack.size + 2 * buffer_sz
    image_offset = [substack.z - buffer_sz, substack.y - buffer_sz, substack.x - buffer_sz]
    full_size = volume.shape
    image = np.zeros((image_sz, image_sz, image_sz), 'uint8')
    lo = np.maximum(image_offset, 0)
    hi = np.minimum(np.asarray(image_offset) + image_sz, full_size)
    if lo[0] > hi[0] or lo[1] > hi[1] or lo[2] > hi[2]:
        return (None, substack)
    image[(lo[0] - image_offset[0]):(hi[0] - image_offset[0]),
          (lo[1] - image_offset[1]):(hi[1] - image_offset[1]),
          (lo[2] - image_offset[2]):(hi[2] - image_offset[2])] = volume[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
    im_raw_mn, im_raw_std = np.mean(image), np.std(image)
    idx = ((image < 200) & (image > 1))
    if np.sum(idx) > 0:
        im_flt_mn, im_flt_std = np.mean(image[idx]), np.std(image[idx])
    else:
        im_flt_mn, im_flt_std = image_normalize[0], image_normalize[1]
    global_frac = 1. if len(image_normalize) < 3 else image_normalize[2]
    mn_use = global_frac * image_normalize[0] + (1 - global_frac) * im_flt_mn
    image = (image.astype('float32') - mn_use) / image_normalize[1]
    norm_fn = '%s/%d_%d_%d_%d.txt' % (norm_dir, substack.size, substack.z, substack.y, substack.x)
    with open(norm_fn, 'w') as f_out:
        f_out.write('%d,%d,%d,%d,%d,%g,%g,%g,%g,%g,%g,%g,%g\n' %
                    (substack.size, buffer_sz, substack.z, substack.y, substack.x,
                     image_normalize[0], image_normalize[1], global_frac, mn_use,
                     im_flt_mn, im_flt_std, im_raw_mn, im_raw_std))
    return (image, substack)


def fri_postprocess(pred, working_dir, obj_min_dist, smoothing_sigma, substack, buffer_sz, thd):
    """voxel2obj of one substack's prediction with the substack's (x,y,z) offset and buffer; the result is
    pickled to fri_filename (fplobjdetect.py:1114-1151).  `pred` may be a numpy array or a CUDA tensor
    (the map FplNetwork.infer_device left in HBM).  Returns the pickle's file name."""
    ff = fri_filename(working_dir, substack)
    if pred is None:
        out = {'locs': np.zeros((0, 3)), 'conf': np.zeros(0)}
    else:
        out = voxel2obj(pred, obj_min_dist, smoothing_sigma,
                        (substack.x - buffer_sz, substack.y - buffer_sz, substack.z - buffer_sz),
                        buffer_sz, thd)
    tmp = '%s.tmp.%d' % (ff, os.getpid())          # atomic: a resuming / neighbouring rank never reads half a pickle
    with open(tmp, 'wb') as f_out:
        pickle.dump(out, f_out)
    os.replace(tmp, ff)
    return ff


def full_roi_inference(data_source, dvid_uuid, dvid_roi, network, thd, working_dir, image_normalize,
                       obj_min_dist=27, smoothing_sigma=5, buffer_sz=35, partition_size=16,
                       local_cache_dir=None, roi_force_file=False, instance_name='grayscale',
                       dvid_seg_info=None):
    """Predictions of a trained network inside an ROI, substack by substack, cached on disk so that a
    second call resumes (flypylib/fplobjdetect.py:841-986; parameters as there).

    ``data_source``: (Z,Y,X) uint8 array-like (see the section comment); ``dvid_roi``: ROI text file
    (``size,z,y,x`` per line); ``dvid_uuid``, ``partition_size``, ``local_cache_dir``, ``roi_force_file``,
    ``instance_name`` are accepted for signature compatibility and unused; ``dvid_seg_info`` must be None
    (segmentation-aware suppression is out of scope).  Returns ``{'locs','conf'}`` over all substacks in
    ROI-file order and writes it to ``<working_dir>/all.p``.

    With ``torch.distributed`` initialised (one process per GPU) rank r processes the pending substacks
    r, r+world, ...; per-substack pickles are written by the rank that computed them and every rank
    returns the complete result (all-gather of the detection lists)."""
    if dvid_seg_info is not None:
        raise NotImplementedError("segmentation-aware suppression is not part of the B200 hot path")
    if isinstance(data_source, str):
        raise NotImplementedError("DVID / DICED / N5 volume services are out of scope; pass the volume as an "
                                  "array-like (numpy.memmap works for volumes larger than host memory)")
    os.makedirs(working_dir, exist_ok=True)
    norm_dir = '%s/norm' % working_dir
    os.makedirs(norm_dir, exist_ok=True)
    roi = roi_from_txt(dvid_roi)

    world, rank = 1, 0
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
    except ImportError:
        dist = None

    # which substacks are already on disk is decided ONCE, by rank 0, before any rank writes: ranks that scanned on
    # their own could see each other's fresh pickles, deal the rest differently and leave substacks unprocessed
    done_idx = None
    if rank == 0:
        done_idx = [si for si, rr in enumerate(roi[0]) if os.path.isfile(fri_filename(working_dir, rr))]
    if world > 1:
        box = [done_idx]
        dist.broadcast_object_list(box, src=0)
        done_idx = box[0]
    done_set = set(done_idx)
    results = {}                    # substack index -> (locs, conf)
    pending = []
    for si, rr in enumerate(roi[0]):
        if si in done_set:
            with open(fri_filename(working_dir, rr), 'rb') as f_in:
                obj = pickle.load(f_in)
            results[si] = (obj['locs'], obj['conf'])
            continue
        pending.append((si, szyx(rr.size, rr.z, rr.y, rr.x)))
    if rank == 0:
        print('already processed: %d' % len(results))
        print('to process: %d' % len(pending))

    n_done = 0
    for pi, (si, ss) in enumerate(pending):
        if pi % world != rank:
            continue
        image, _ = fri_get_image([ss, image_normalize, buffer_sz, norm_dir], data_source)
        pred = None
        if image is not None:
            # keep the map on the device when the network offers it (no 4 B/voxel round trip over PCIe)
            if hasattr(network, 'infer_device'):
                import torch
                pred = network.infer_device(torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).cuda())
            else:
                pred = network.infer(image)
        ff = fri_postprocess(pred, working_dir, obj_min_dist, smoothing_sigma, ss, buffer_sz, thd)
        with open(ff, 'rb') as f_in:
            obj = pickle.load(f_in)
        results[si] = (obj['locs'], obj['conf'])
        n_done += 1
        if rank == 0:
            sys.stdout.write('\r%d' % n_done)
            sys.stdout.flush()

    if world > 1:
        mine = [(si,) + results[si] for pi, (si, _) in enumerate(pending) if pi % world == rank]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        for part in gathered:
            for si, l, c in part:
                results[si] = (l, c)

    missing = [si for si in range(len(roi[0])) if si not in results]
    if missing:
        raise RuntimeError("full_roi_inference: substacks %s were not processed" % missing[:8])
    order = sorted(results)
    locs = np.concatenate([np.asarray(results[si][0]).reshape(-1, 3) for si in order]) if order else np.zeros((0, 3))
    conf = np.concatenate([np.asarray(results[si][1]).reshape(-1) for si in order]) if order else np.zeros(0)
    obj = {'locs': locs, 'conf': conf}
    if rank == 0:
        with open('%s/all.p' % working_dir, 'wb') as f_out:
            pickle.dump(obj, f_out)
    return obj


# ---------------------------------------------------------------------------------------------
# evaluate_substacks (flypylib/fplobjdetect.py:459-512): infer + voxel2obj + PR curve per substack.
# The reference forks one process per substack for the post-processing; a CUDA context does not
# survive fork and voxel2obj is now faster than the fork, so the work runs inline.
# ---------------------------------------------------------------------------------------------
def _get_labels(seg, tt):
    """Segment label under every detection (fplobjdetect.py:459-461); locs are (x,y,z), seg is (Z,Y,X)."""
    tt_ind = tt['locs'].astype(int)
    return seg[tt_ind[:, 2], tt_ind[:, 1], tt_ind[:, 0]]


def evaluate_substacks(network, substacks, thds, obj_min_dist=27, smoothing_sigma=5, volume_offset=(0, 0, 0),
                       buffer_sz=5, allow_mult=False):
    """Precision/recall of ``network`` on annotated substacks (fplobjdetect.py:484-512).

    ``substacks``: sequence of ``(image, groundtruth_json[, seg])`` -- image as accepted by
    ``network.infer`` (array), ground truth as a json file (or json text) in either wire format, optional
    segmentation array (Z,Y,X) for label-constrained matching (the reference takes an h5 path; h5py is not
    available).  Returns ``(aggregate_pr(results), results)`` with one PR_Result per substack, in order."""
    from . import fplsynapses
    results = []
    for ss in substacks:
        image = ss[0]
        if hasattr(network, 'infer_device') and not isinstance(image, str):
            import torch
            dev_img = image if isinstance(image, torch.Tensor) else \
                torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32))
            pred = network.infer_device(dev_img.cuda())
            shape = tuple(pred.shape)
        else:
            pred = network.infer(image)
            shape = pred.shape
        out = voxel2obj(pred, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz)
        gt = fplsynapses.load_from_json(ss[1], shape, buffer_sz)
        lbls_pd = lbls_gt = None
        if len(ss) >= 3 and ss[2] is not None:
            seg = np.asarray(ss[2])
            lbls_pd, lbls_gt = _get_labels(seg, out), _get_labels(seg, gt)
        results.append(obj_pr_curve(out, gt, obj_min_dist, thds, lbls_pd, lbls_gt, allow_mult=allow_mult))
    return aggregate_pr(results), results


# ---------------------------------------------------------------------------------------------
# gen_batches (flypylib/fplobjdetect.py:27-130): the patch generator that feeds FplNetwork.train
# (BASELINE config 5).  Host-side numpy; same sampling and augmentation, same order of np.random calls
# (so a seeded run reproduces the reference's batches bit for bit).  h5py is not available: every
# training volume is given as arrays instead of hdf5 file names.
# ---------------------------------------------------------------------------------------------
def gen_batches(train_data, context_sz, batch_sz, is_mask=False):
    """generator that yields ``(data, labels)`` training batches for ``fit_generator``.

    ``train_data``: sequence of ``(image, labels, mask)`` arrays of equal 3-D shape (the reference takes
    ``(image.h5, prefix)`` and reads ``<prefix>labels.h5`` / ``<prefix>mask.h5``).  ``context_sz``: patch
    size; ``batch_sz``: examples per batch, alternating label 0 / label 1 positions drawn with replacement
    from the masked-in voxels at least half a patch away from the faces.  Augmentation per example:
    rot90 in the last two axes (0-3 times), flip of the last axis, flip of the first axis.
    ``data`` is float32 ``(batch,) + context_sz + (1,)``; ``labels`` uint8 ``(batch,1,1,1,1)`` or, with
    ``is_mask``, ``(batch,6,6,6,1)`` label patches in which masked-out voxels carry label 2.  As in the
    reference the two arrays are re-used between yields."""
    n_per_class = int(round(batch_sz / 2))
    context_rr = tuple(int(round(cc / 2)) for cc in context_sz)
    ims, lls, mms, locs = [], [], [], []
    for tr in train_data:
        ims.append(np.asarray(tr[0]))
        lls.append(np.array(tr[1]))
        mm = np.array(tr[2])
        mm[:context_rr[0], :, :] = 0; mm[:, :context_rr[1], :] = 0; mm[:, :, :context_rr[2]] = 0
        mm[-context_rr[0]:, :, :] = 0; mm[:, -context_rr[1]:, :] = 0; mm[:, :, -context_rr[2]:] = 0
        mms.append(mm)
        locs.append([((lls[-1] == cc) & (mm == 1)).nonzero() for cc in range(2)])
        if is_mask:
            lls[-1][(mm == 0).nonzero()] = 2
    train_idx, n_train = 0, len(train_data)
    data = np.zeros((batch_sz, context_sz[0], context_sz[1], context_sz[2], 1), dtype='float32')
    labels = np.zeros((batch_sz, 6, 6, 6, 1) if is_mask else (batch_sz, 1, 1, 1, 1), dtype='uint8')
    while True:
        im, ll = ims[train_idx], lls[train_idx]
        for cc in range(2):
            pos = locs[train_idx][cc]
            if len(pos[0]) == 0:            # class absent in this volume: keep the previous examples
                continue
            pick = np.random.choice(len(pos[0]), n_per_class, True)
            for ii in range(n_per_class):
                xx, yy, zz = pos[0][pick[ii]], pos[1][pick[ii]], pos[2][pick[ii]]
                ex = ii * 2 + cc
                data[ex, :, :, :, 0] = im[xx - context_rr[0]:xx + context_rr[0], yy - context_rr[1]:yy + context_rr[1],
                                          zz - context_rr[2]:zz + context_rr[2]]
                if is_mask:
                    labels[ex, :, :, :, 0] = ll[xx - 3:xx + 3, yy - 3:yy + 3, zz - 3:zz + 3]
                else:
                    labels[ex, 0] = ll[xx, yy, zz]
        aug_rot = np.floor(4 * np.random.rand(batch_sz))
        aug_ref = np.floor(2 * np.random.rand(batch_sz))
        aug_fpz = np.floor(2 * np.random.rand(batch_sz))
        for ii in range(batch_sz):
            if aug_rot[ii]:
                data[ii, :, :, :, 0] = np.rot90(data[ii, :, :, :, 0], int(aug_rot[ii]), (1, 2))
                if is_mask:
                    labels[ii, :, :, :, 0] = np.rot90(labels[ii, :, :, :, 0], int(aug_rot[ii]), (1, 2))
            if aug_ref[ii]:
                data[ii, :, :, :, 0] = np.flip(data[ii, :, :, :, 0], 2)
                if is_mask:
                    labels[ii, :, :, :, 0] = np.flip(labels[ii, :, :, :, 0], 2)
            if aug_fpz[ii]:
                data[ii, :, :, :, 0] = np.flip(data[ii, :, :, :, 0], 0)
                if is_mask:
                    labels[ii, :, :, :, 0] = np.flip(labels[ii, :, :, :, 0], 0)
        yield data, labels
        train_idx = (train_idx + 1) % n_train


def get_out_sz(in_sz):
    """output edge of a u-net style net for an input sub-volume edge (flypylib/fplobjdetect.py:514-522)"""
    import math
    bottleneck_sz = int(math.floor(math.floor((in_sz - 2) / 2) - 2) / 2)
    return (bottleneck_sz * 2 - 2) * 2 - 2
