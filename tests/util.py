"""Shared parity metrics (BASELINE.json north_star tolerances) and a JSON-lines parity report.

Relevance maps are consumed after abs-max normalisation (reference: innvestigate/utils/visualizations.py:36-54
`project`, models/model.py:1676-1680 `hp /= max|hp|`), and their values span many orders of magnitude, so the
per-pixel / per-feature relative error is measured against the map's absolute maximum:

    linf_rel = max|got - ref| / max|ref|        l2_rel = ||got - ref||_2 / ||ref||_2

`floor_rel` (denominator |ref| + 1e-3 max|ref|) is also recorded in the report; it is dominated by near-zero
pixels and is informative only for the rules whose arithmetic has no cancellation (alpha-beta family).
"""
import json
import os

import numpy as np

REL_TOL = 1e-3      # per-pixel / per-feature error relative to the map's abs-max, and relative L2
SUM_TOL = 1e-4      # relevance-conservation sums (relative to the sum of |R|)

_REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def _f64(a):
    return np.asarray(a, dtype=np.float64)


def linf_rel(got, ref):
    got, ref = _f64(got), _f64(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got)), "non-finite values in result"
    scale = np.max(np.abs(ref))
    return float(np.max(np.abs(got - ref)) / scale) if scale > 0 else float(np.max(np.abs(got)))


def l2_rel(got, ref):
    got, ref = _f64(got), _f64(ref)
    n = np.linalg.norm(ref.ravel())
    return float(np.linalg.norm((got - ref).ravel()) / n) if n > 0 else float(np.linalg.norm(got.ravel()))


def floor_rel(got, ref, floor=1e-3):
    got, ref = _f64(got), _f64(ref)
    scale = np.max(np.abs(ref))
    if scale == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + floor * scale)))


def sum_err(got, ref, mass=None):
    """Relevance-conservation error: |sum(got) - sum(ref)| relative to the relevance mass sum|ref|.  `mass` (optional):
    the mass that was fed in, sum|R_head| -- a synthetic mixed-sign head can cancel to a map whose own mass is hundreds of
    times smaller than what was propagated, and the rounding of the propagation is relative to the latter; the
    denominator is then the larger of the two masses."""
    got, ref = _f64(got), _f64(ref)
    den = np.abs(ref).sum()
    if mass is not None:
        den = max(den, float(mass))
    return float(abs(got.sum() - ref.sum()) / (den + 1e-300))


def topk_cells(R, k=10, cell=16):
    """Indices of the k most relevant grid regions: R [H, W, C] summed over cell x cell blocks and channels."""
    R = _f64(R)
    H, W = R.shape[0], R.shape[1]
    g = R.reshape(H // cell, cell, W // cell, cell, -1).sum(axis=(1, 3, 4))
    order = np.argsort(-g.reshape(-1), kind="stable")
    return [int(i) for i in order[:min(k, order.size)]]


def cell_sums(R, cell=16):
    R = _f64(R)
    H, W = R.shape[0], R.shape[1]
    return R.reshape(H // cell, cell, W // cell, cell, -1).sum(axis=(1, 3, 4)).reshape(-1)


def same_topk(got, ref, k=10, cell=16, tol=REL_TOL, sums=None):
    """"Identical top-k relevant grid regions" up to ties inside the stated tolerance: the two rankings must agree rank
    by rank, except where the reference's own scores of the two cells involved differ by less than tol * max|score| (a
    difference the per-pixel tolerance itself allows to go either way).  sums: callable giving the per-cell scores."""
    f = sums or (lambda a: cell_sums(a, cell))
    g, r = f(got), f(ref)
    og = np.argsort(-g, kind="stable")[:min(k, g.size)]
    orf = np.argsort(-r, kind="stable")[:min(k, r.size)]
    scale = np.max(np.abs(r))
    for a, b in zip(og, orf):
        if a != b and abs(r[a] - r[b]) > tol * scale:
            return False
    return True


def topk_features(R, k=10):
    """Top-k grid cells of a feature-level relevance map R [L, D] or [h, w, D] (summed over D)."""
    g = _f64(R).sum(axis=-1).reshape(-1)
    return [int(i) for i in np.argsort(-g, kind="stable")[:min(k, g.size)]]


def record(what, got, ref, **extra):
    m = {"what": what, "linf_rel": linf_rel(got, ref), "l2_rel": l2_rel(got, ref), "floor_rel": floor_rel(got, ref),
         "sum_err": sum_err(got, ref)}
    m.update(extra)
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        with open(_REPORT, "a") as f:
            f.write(json.dumps(m) + "\n")
    except OSError:
        pass
    return m


def assert_parity(got, ref, what="", rel_tol=REL_TOL, sum_tol=SUM_TOL, **extra):
    m = record(what, got, ref, rel_tol=rel_tol, **extra)
    assert m["linf_rel"] <= rel_tol, "%s: linf-rel %.3e > %.1e (l2 %.3e)" % (what, m["linf_rel"], rel_tol, m["l2_rel"])
    assert m["l2_rel"] <= rel_tol, "%s: l2-rel %.3e > %.1e" % (what, m["l2_rel"], rel_tol)
    if sum_tol is not None:
        assert m["sum_err"] <= sum_tol, "%s: sum-err %.3e > %.1e" % (what, m["sum_err"], sum_tol)
    return m
