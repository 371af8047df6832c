"""TEST INFRASTRUCTURE ONLY -- exact solver for the matching program of the reference's ``obj_match``
(flypylib/fplobjdetect.py:259-318), used in place of PuLP (absent from this image) when the
reference's ``obj_pr`` / ``obj_pr_curve`` are run unmodified to generate golden vectors, and as the
checker of ``flypylib_b200.fplobjdetect.obj_match`` in tests/.

The program: binary x_ij for every pair with dists[i,j] < 0; minimise sum dists[i,j] x_ij subject to
sum_i x_ij <= 1 for every ground truth j and (unless allow_mult) sum_j x_ij <= 1 for every
prediction i.  Solved by exhaustive depth-first enumeration over ground-truth columns -- exponential,
for small cases only.  **Parity unpinned against PuLP itself** (not installable); the optimum VALUE
of an integer program is solver independent, which is what the comparisons use (total cost and
number of matches; the reference's precision/recall depend on the match count only).
"""
import numpy as np


def obj_match(dists, allow_mult=False):
    d = np.asarray(dists, dtype=np.float64)
    n_pred, n_gt = d.shape
    cand = [np.nonzero(d[:, j] < 0)[0] for j in range(n_gt)]
    best = {"cost": 0.0, "pick": [-1] * n_gt}
    pick = [-1] * n_gt
    used = np.zeros(n_pred, dtype=bool)
    # optimistic bound: every remaining column takes its cheapest admissible row
    col_min = np.array([d[c, j].min() if c.size else 0.0 for j, c in enumerate(cand)])
    tail = np.concatenate([np.cumsum(col_min[::-1])[::-1], [0.0]])

    def rec(j, cost):
        if j == n_gt:
            if cost < best["cost"]:
                best["cost"], best["pick"] = cost, list(pick)
            return
        if cost + tail[j] >= best["cost"]:          # cannot strictly improve on the incumbent
            return
        for i in cand[j]:
            if allow_mult or not used[i]:
                used[i] = True; pick[j] = i
                rec(j + 1, cost + d[i, j])
                used[i] = False; pick[j] = -1
        rec(j + 1, cost)

    rec(0, 0.0)
    out = np.zeros((n_pred, n_gt), dtype=bool)
    for j, i in enumerate(best["pick"]):
        if i >= 0:
            out[i, j] = True
    return out
