"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of flypylib's ``voxel2obj``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product (``flypylib_b200``) never does.

Reference path restated here (file:line relative to /root/reference):

* ``flypylib/fplobjdetect.py:132-257``  voxel2obj  (seg=None branch)
* ``flypylib/fplutils.py:9-22``         to3d, set_filter (ball mask ``sqrt(dz^2+dy^2+dx^2) <= r``)

Third-party arithmetic the reference delegates to, restated from the libraries installed
in this image (the reference's conda recipe does not pin them, conda-recipe/meta.yaml:15-22):

* ``scipy.ndimage.gaussian_filter(pred, sigma, truncate=2.0)`` (SciPy 1.18.1,
  ``ndimage/_filters.py`` gaussian_filter -> gaussian_filter1d -> correlate1d, C routine
  ``NI_Correlate1D`` symmetric-kernel branch): per axis 0,1,2; double line buffer; mode
  'reflect'; ``tmp = x[c]*w[0]; for j=-R..-1: tmp += (x[c+j] + x[c-j]) * w[j]`` with
  separate (non fused) multiply and add; result rounded to float32 after every axis.
* ``np.percentile(pred, 97)`` on a float32 array (NumPy 2.3.5, ``lib/_function_base_impl.py``
  percentile -> _quantile, method 'linear'): q = float32(97)/float32(100);
  virtual index = float32(n-1) * q in float32; lerp in float32.

Parity pin: ``tests/golden/voxel2obj_*.npz`` hold outputs of the *unmodified* reference
function (run in the build container through ``oracle/ref_loader.py`` by
``tests/golden/make_golden.py``); ``tests/test_oracle_voxel2obj.py`` checks every function
here bit-for-bit against them.

Two implementations of every stage are provided:
  * ``impl='numpy'``  literal, vectorised numpy; the greedy loop is the reference's
    O(K*C) formulation (small cases);
  * ``impl='c'``      ``oracle/voxel2obj_c.c`` (gcc, -ffp-contract=off, OpenMP over lines)
    with the greedy loop restated as "visit candidates in (value desc, index asc) order,
    skip invalid ones" -- the same selection, O(C log C + K*ball).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------- helpers
def to3d(v):
    """fplutils.py:9-12 -- a scalar becomes a 3-tuple."""
    if np.size(v) == 1:
        v = (v, v, v)
    return v


def ball_mask(r):
    """fplutils.py:14-22 set_filter(r): (2r+1)^3 bool, True where sqrt(d2) <= r."""
    ax = np.arange(-r, r + 1)
    d2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
    return np.sqrt(d2) <= r


def gaussian_weights(sigma, truncate=2.0):
    """SciPy _gaussian_kernel1d(sigma, 0, lw)[::-1] with lw = int(truncate*sigma+0.5)."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return phi[::-1].copy(), lw


def _reflect_index(idx, n):
    """SciPy 'reflect' (d c b a | a b c d | d c b a) index map for arbitrary overshoot."""
    period = 2 * n
    m = np.mod(idx, period)
    return np.where(m >= n, period - 1 - m, m)


def correlate1d_symmetric(a32, weights, lw, axis):
    """One gaussian_filter1d pass restated: float32 in -> double accumulate -> float32 out."""
    n = a32.shape[axis]
    a = np.moveaxis(a32, axis, 0).astype(np.float64)
    ext_idx = _reflect_index(np.arange(-lw, n + lw), n)
    ext = a[ext_idx]                                   # (n + 2 lw, ...)
    fw = weights[lw:]                                  # fw[0] centre ... symmetric so fw[j]==w[lw-j]
    centre = ext[lw:lw + n]
    tmp = centre * weights[lw]
    for jj in range(-lw, 0):
        pair = ext[lw + jj:lw + jj + n] + ext[lw - jj:lw - jj + n]
        tmp = tmp + pair * weights[lw + jj]
    out = tmp.astype(np.float32)
    return np.moveaxis(out, 0, axis)


def gaussian_filter_f32(a32, sigma, truncate=2.0):
    """scipy.ndimage.gaussian_filter(a32, sigma, truncate=truncate) for a float32 3-D array."""
    out = np.ascontiguousarray(a32, dtype=np.float32)
    if not float(sigma) > 1e-15:
        return out.copy()
    w, lw = gaussian_weights(sigma, truncate)
    for axis in range(out.ndim):
        out = correlate1d_symmetric(out, w, lw, axis)
    return np.ascontiguousarray(out)


def percentile_plan(n, q=97):
    """Order statistics + weight that np.percentile(float32 array of n values, q) uses.

    Returns (prev_index, next_index, gamma float32).  NumPy 2.3.5 _quantile, 'linear'.
    """
    qf = np.true_divide(q, np.float32(100))            # float32 scalar (q is a weak python number)
    vi = np.asanyarray((n - 1) * qf)                     # float32
    prev = np.asanyarray(np.floor(vi))
    nxt = np.asanyarray(prev + 1)
    if vi >= n - 1:
        prev = np.asanyarray(np.float32(-1)); nxt = np.asanyarray(np.float32(-1))
    if vi < 0:
        prev = np.asanyarray(np.float32(0)); nxt = np.asanyarray(np.float32(0))
    gamma = np.asanyarray(vi - prev, dtype=vi.dtype)
    pi = int(prev.astype(np.intp)); ni = int(nxt.astype(np.intp))
    if pi < 0:
        pi += n
    if ni < 0:
        ni += n
    return pi, ni, gamma[()]


def percentile_lerp(lo, hi, gamma):
    """NumPy _lerp on float32 scalars."""
    lo = np.float32(lo); hi = np.float32(hi); gamma = np.float32(gamma)
    d = hi - lo
    res = lo + d * gamma
    if gamma >= 0.5:
        res = hi - d * (1 - gamma)
    return np.float32(res)


def percentile_f32(a32, q=97):
    """np.percentile(a32, q) restated through order statistics (float32 input)."""
    flat = np.ascontiguousarray(a32, dtype=np.float32).ravel()
    n = flat.size
    if np.isnan(flat).any():
        return np.float32(np.nan)
    pi, ni, gamma = percentile_plan(n, q)
    part = np.partition(flat, sorted({pi, ni}))
    return percentile_lerp(part[pi], part[ni], gamma)


# --------------------------------------------------------------------------- stages
def smooth_padded(pred, r, sigma):
    """fplobjdetect.py:158-175: pad r zeros, gaussian, zero the r-wide border."""
    p = np.pad(np.asarray(pred), r, 'constant')
    if p.dtype != np.float32:
        raise TypeError("oracle restates the float32 path only")
    s = gaussian_filter_f32(p, sigma)
    if r > 0:
        s[:r] = 0; s[:, :r] = 0; s[:, :, :r] = 0
        s[-r:] = 0; s[:, -r:] = 0; s[:, :, -r:] = 0
    else:
        s[...] = 0        # x[-0:] = 0 in the reference clears everything
    return s


def threshold(s, thd):
    """fplobjdetect.py:183: max(97th percentile, thd)."""
    return np.maximum(percentile_f32(s, 97), thd)


def greedy_nms_literal(s, thresh, r):
    """fplobjdetect.py:184-231 restated literally (argmax / invalidate ball / drop)."""
    cand = np.flatnonzero(s.ravel() > thresh)
    flat = s.ravel()
    keep = ~ball_mask(r)
    valid = np.ones(s.shape, dtype=bool)
    vflat = valid.reshape(-1)
    rows = []
    while cand.size:
        vals = flat[cand]
        j = int(np.argmax(vals))
        if vals[j] <= 0:
            break
        z, y, x = np.unravel_index(cand[j], s.shape)
        rows.append((x, y, z, vals[j]))
        valid[z - r:z + r + 1, y - r:y + r + 1, x - r:x + r + 1] &= keep
        cand = cand[vflat[cand]]
    return rows


def greedy_nms_sorted(s, thresh, r):
    """Same selection as greedy_nms_literal: walk candidates in (value desc, index asc)."""
    flat = s.ravel()
    cand = np.flatnonzero(flat > thresh)
    vals = flat[cand]
    order = np.lexsort((cand, -vals.astype(np.float64)))
    keep = ~ball_mask(r)
    valid = np.ones(s.shape, dtype=bool)
    vflat = valid.reshape(-1)
    rows = []
    for k in order:
        i = cand[k]
        if not vflat[i]:
            continue
        if vals[k] <= 0:
            break
        z, y, x = np.unravel_index(i, s.shape)
        rows.append((x, y, z, vals[k]))
        valid[z - r:z + r + 1, y - r:y + r + 1, x - r:x + r + 1] &= keep
    return rows


def greedy_nms_seg(s, thresh, r, seg_padded, seg_dilate=None, seg_force=None):
    """fplobjdetect.py:187-231 with a segmentation (:192-195, :213-224), restated literally: the suppression mask of a
    selected point is ball AND dilate(seg cube == seg[point]) (OR the forced inner ball)."""
    from scipy import ndimage
    flat = s.ravel()
    cand = np.flatnonzero(flat > thresh)
    dist_flt = ~ball_mask(r)
    if seg_force:
        cn = ~np.pad(ball_mask(seg_force), r - seg_force, 'constant')
    valid = np.ones(s.shape, dtype=bool)
    vflat = valid.reshape(-1)
    rows = []
    while cand.size:
        vals = flat[cand]
        j = int(np.argmax(vals))
        if vals[j] <= 0:
            break
        z, y, x = np.unravel_index(cand[j], s.shape)
        rows.append((x, y, z, vals[j]))
        keep = dist_flt
        if seg_padded is not None:
            m = seg_padded[z - r:z + r + 1, y - r:y + r + 1, x - r:x + r + 1] == seg_padded[z, y, x]
            if seg_dilate is not None:
                m = ndimage.binary_dilation(m, iterations=seg_dilate)
            keep = np.logical_not(m) | dist_flt
            if seg_force:
                keep = keep & cn
        valid[z - r:z + r + 1, y - r:y + r + 1, x - r:x + r + 1] &= keep
        cand = cand[vflat[cand]]
    return rows


def voxel2obj_seg(pred, obj_min_dist, smoothing_sigma, volume_offset=(0, 0, 0), buffer_sz=0, thd=0,
                  seg=None, seg_dilate=None, seg_sz_thd=None, seg_force=None):
    """Oracle for the segmentation-aware call of fplobjdetect.voxel2obj (:161-165, 177-181, 192-195, 213-224)."""
    r = obj_min_dist
    pred = np.asarray(pred)
    s = smooth_padded(pred, r, smoothing_sigma)
    segp = np.pad(np.asarray(seg), r, 'constant') if seg is not None else None
    if seg_sz_thd is not None:
        ids, counts = np.unique(segp, return_counts=True)
        for i, c in zip(ids, counts):
            if c < seg_sz_thd:
                s[segp == i] = False
    t = threshold(s, thd)
    rows = greedy_nms_seg(s, t, r, segp, seg_dilate, seg_force)
    return finish(rows, r, pred.shape, buffer_sz, volume_offset)


def finish(rows, r, pred_sz, buffer_sz, volume_offset):
    """fplobjdetect.py:233-257: to (K,4) float64, un-pad, buffer crop, offset, split."""
    b = to3d(buffer_sz)
    if rows:
        a = np.asarray([[float(x), float(y), float(z), float(v)] for x, y, z, v in rows],
                       dtype=np.float64)
    else:
        a = np.zeros((0, 4))
    a[:, :3] -= r
    lo = np.array([b[0], b[1], b[2]], dtype=np.float64)
    hi = np.array([pred_sz[2] - b[0], pred_sz[1] - b[1], pred_sz[0] - b[2]], dtype=np.float64)
    ok = np.all(a[:, :3] >= lo, axis=1) & np.all(a[:, :3] < hi, axis=1)
    a = a[ok]
    a = a + np.array([tuple(volume_offset) + (0,)])
    return {'locs': a[:, :3], 'conf': a[:, 3]}


# --------------------------------------------------------------------------- C helper
_clib = None


def build_c(force=False):
    """Compile oracle/voxel2obj_c.c -> oracle/_build/libfploracle.so (gcc, no FMA contraction)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libfploracle.so")
    src = os.path.join(_HERE, "voxel2obj_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               "-o", so, src, "-lm"])
    return so


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(build_c())
        lib.fpl_oracle_smooth_padded.restype = ctypes.c_int
        lib.fpl_oracle_smooth_padded.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        lib.fpl_oracle_greedy.restype = ctypes.c_int64
        lib.fpl_oracle_greedy.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
            ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _clib = lib
    return _clib


def smooth_padded_c(pred, r, sigma, threads=0):
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    Z, Y, X = pred.shape
    out = np.zeros((Z + 2 * r, Y + 2 * r, X + 2 * r), dtype=np.float32)
    if float(sigma) > 1e-15:
        w, lw = gaussian_weights(sigma)
        w = np.ascontiguousarray(w, dtype=np.float64)
    else:
        w, lw = np.ones(1), -1
    rc = _c().fpl_oracle_smooth_padded(pred.ctypes.data, Z, Y, X, r, w.ctypes.data, lw,
                                      out.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError("fpl_oracle_smooth_padded failed: %d" % rc)
    return out


def greedy_nms_c(s, thresh, r, max_out=None):
    s = np.ascontiguousarray(s, dtype=np.float32)
    Z, Y, X = s.shape
    if max_out is None:
        max_out = max(1024, s.size // max(1, (r + 1) ** 3) * 8 + 1024)
    idx = np.zeros(max_out, dtype=np.int64)
    val = np.zeros(max_out, dtype=np.float32)
    k = _c().fpl_oracle_greedy(s.ctypes.data, Z, Y, X, r, float(thresh), idx.ctypes.data,
                               val.ctypes.data, max_out)
    if k < 0:
        raise RuntimeError("fpl_oracle_greedy failed: %d" % k)
    rows = []
    for i in range(k):
        z, rem = divmod(int(idx[i]), Y * X)
        y, x = divmod(rem, X)
        rows.append((x, y, z, val[i]))
    return rows


# --------------------------------------------------------------------------- entry point
def voxel2obj(pred, obj_min_dist, smoothing_sigma, volume_offset=(0, 0, 0), buffer_sz=0, thd=0,
              impl='numpy', return_intermediates=False):
    """Oracle for fplobjdetect.voxel2obj(pred, r, sigma, volume_offset, buffer_sz, thd), seg=None."""
    r = obj_min_dist
    pred = np.asarray(pred)
    if impl == 'c':
        s = smooth_padded_c(pred, r, smoothing_sigma)
    else:
        s = smooth_padded(pred, r, smoothing_sigma)
    t = threshold(s, thd)
    if impl == 'c':
        rows = greedy_nms_c(s, t, r)
    elif impl == 'numpy-literal':
        rows = greedy_nms_literal(s, t, r)
    else:
        rows = greedy_nms_sorted(s, t, r)
    out = finish(rows, r, pred.shape, buffer_sz, volume_offset)
    if return_intermediates:
        return out, s, t
    return out
