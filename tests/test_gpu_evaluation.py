"""Device evaluation reductions against the numpy restatement of the reference helpers."""
import numpy as np
import pytest

from lrp_imagecaptioning_b200 import evaluation as EV
from oracle import evaluation_ref as ER

pytestmark = pytest.mark.gpu


def _maps(n, hw, seed, sparse=False):
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((n, hw, hw, 3)).astype(np.float32) * rng.random((n, 1, 1, 1)).astype(np.float32)
    if sparse:
        m *= (rng.random((n, hw, hw, 1)) > 0.7)
    m[n - 1] = 0.0          # an all-zero map: project returns zeros
    return m


@pytest.mark.parametrize("mode,shift", [("mean", False), ("mean", True), ("negative", True), ("positive", False)])
@pytest.mark.parametrize("hw", [32, 224])
def test_heatmaps_match_reference_helpers(mode, shift, hw):
    m = _maps(5, hw, hw)
    got, means = EV.heatmaps(m, mode, shift_negative=shift, want_means=True)
    got = got.cpu().numpy()
    for i in range(m.shape[0]):
        ref = ER.heatmap(m[i], mode, shift)
        assert np.abs(got[i] - ref).max() <= 1e-6, (mode, i)
        assert abs(means[i] - np.mean(ref)) <= 1e-6


@pytest.mark.parametrize("kind", ["max", "ave"])
def test_pooled_heatmaps(kind):
    m = _maps(4, 224, 7)
    got = EV.heatmaps(m, "mean", window=16, pooling=kind).cpu().numpy()
    assert got.shape == (4, 14, 14)
    for i in range(4):
        assert np.abs(got[i] - ER.pooled_heatmap(m[i], 16, kind)).max() <= 2e-6


def test_bbox_correctness_all_thresholds():
    m = _maps(6, 224, 11, sparse=True)
    heat = EV.heatmaps(m, "negative", shift_negative=True)
    h = heat.cpu().numpy()
    rng = np.random.default_rng(3)
    boxes = []
    for i in range(6):
        for _ in range(3):
            x0, y0 = rng.integers(0, 150, 2)
            boxes.append((i, x0, y0, x0 + rng.integers(1, 74), y0 + rng.integers(1, 74)))
    boxes.append((0, 0, 0, 224, 224))      # whole image -> 1
    boxes.append((1, 10, 10, 10, 40))      # empty box -> 0
    got = EV.bbox_correctness(heat, boxes)
    assert got.shape == (len(boxes), len(EV.THRESHOLDS))
    for bi, (mi, x0, y0, x1, y1) in enumerate(boxes):
        for ti, th in enumerate(EV.THRESHOLDS):
            ref = ER.overlapped_pixels([x0, y0, x1, y1], h[mi].astype(np.float32), np.float32(th))
            assert abs(got[bi, ti] - ref) <= 2e-6, (bi, ti, got[bi, ti], ref)
    assert np.allclose(got[-2, 0], 1.0) and np.all(got[-1] == 0)
    assert np.all(got[[b[0] == 5 for b in boxes]] == 0)    # the all-zero map


def test_argument_errors():
    m = _maps(2, 32, 0)
    with pytest.raises(ValueError):
        EV.heatmaps(m, "median")
    with pytest.raises(ValueError):
        EV.heatmaps(m, "mean", window=5)
    from lrp_imagecaptioning_b200._lib import LrpcapError
    with pytest.raises(LrpcapError):
        EV.bbox_correctness(EV.heatmaps(m), [(7, 0, 0, 4, 4)])


def test_against_reference_fixture():
    """tests/golden/evaluation.npz holds the outputs of the reference's own evaluation methods."""
    import os
    from oracle.make_golden import evaluation_case
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "evaluation.npz"))
    maps, boxes, thresholds = evaluation_case()
    heat = EV.heatmaps(maps, "negative", shift_negative=True)
    assert np.abs(heat.cpu().numpy() - z["heat_negative"]).max() <= 1e-6
    ratios = EV.bbox_correctness(heat, boxes, thresholds)
    assert np.abs(ratios - z["ratios"]).max() <= 2e-6
    # the reference pools the un-projected channel mean; the device call projects afterwards -> compare after project
    for kind, key in (("max", "pool_max"), ("ave", "pool_ave")):
        got = EV.heatmaps(maps, "mean", window=16, pooling=kind).cpu().numpy()
        for i in range(maps.shape[0]):
            ref = z[key][i] / np.abs(z[key][i]).max()
            assert np.abs(got[i] - ref).max() <= 2e-6
