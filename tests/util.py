"""Shared parity metrics (BASELINE.json north_star tolerances)."""
import numpy as np

REL_TOL = 1e-3      # per-pixel / per-feature relative error
SUM_TOL = 1e-4      # relevance-conservation sums (relative to the sum of |R|)
FLOOR = 1e-3        # rel-err denominator floor, as a fraction of max|ref| (SURVEY.md §7 "Precision vs. tolerance")


def rel_err(got, ref):
    """max over elements of |got - ref| / (|ref| + FLOOR * max|ref|)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got)), "non-finite values in result"
    scale = np.max(np.abs(ref))
    if scale == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + FLOOR * scale)))


def sum_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(abs(got.sum() - ref.sum()) / (np.abs(ref).sum() + 1e-300))


def topk_cells(R, k=10, cell=16):
    """Indices of the k most relevant grid regions: R [H, W, C] pooled over cell x cell blocks and channels."""
    R = np.asarray(R, dtype=np.float64)
    H, W = R.shape[0], R.shape[1]
    g = R.reshape(H // cell, cell, W // cell, cell, -1).sum(axis=(1, 3, 4))
    order = np.argsort(-g.reshape(-1), kind="stable")
    return list(order[:min(k, order.size)])


def topk_features(R, k=10):
    """Top-k grid cells of a feature-level relevance map R [h, w, D] (summed over D)."""
    g = np.asarray(R, dtype=np.float64).sum(axis=-1).reshape(-1)
    return list(np.argsort(-g, kind="stable")[:min(k, g.size)])


def assert_parity(got, ref, what="", rel_tol=REL_TOL, sum_tol=SUM_TOL):
    r = rel_err(got, ref)
    s = sum_err(got, ref)
    assert r <= rel_tol, "%s: rel-err %.3e > %.1e" % (what, r, rel_tol)
    assert s <= sum_tol, "%s: sum-err %.3e > %.1e" % (what, s, sum_tol)
    return r, s
