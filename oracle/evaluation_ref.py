"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's explanation-evaluation helpers (numpy).

Follows, line by line:
  postprocess / _postprocess_images   innvestigate/utils/__init__.py:123-139, evaluate_bbox.py:88-102
  project                             exaimin_word.py:80-89 (no shift), evaluate_bbox.py:60-69 (shift when negative)
  heat map                            exaimin_word.py:96-103 ; evaluate_bbox.py:79-84
  _max_pooling / _ave_pooling         exaimin_word.py:64-77
  _calculate_overlaped_pixels         evaluate_bbox.py:191-208
Pinned against the reference's own methods (run under oracle/refstub.py) by tests/test_oracle_pinning.py."""
import numpy as np


def postprocess(R, color_coding="BGRtoRGB"):
    return R[..., ::-1] if color_coding in ("BGRtoRGB", "RGBtoBGR") else R


def project(x, shift_negative=False):
    absmax = np.max(np.abs(x))
    if absmax == 0:
        return np.zeros(x.shape)
    x = 1.0 * x / absmax
    if shift_negative and np.sum(x < 0):
        x = (x + 1) / 2
    return x


def heatmap(R, mode="mean", shift_negative=False):
    """R (hw, hw, 3) float32 -> (hw, hw)."""
    hm = postprocess(np.asarray(R, dtype=np.float32))
    if mode == "negative":
        hm = np.maximum(-1 * hm, 0)
    elif mode == "positive":
        hm = np.maximum(hm, 0)
    return project(np.mean(hm, axis=-1), shift_negative)


def pool(hp, window=16, kind="max"):
    n = hp.shape[0] // window
    out = np.zeros((n, n))
    for i in range(0, hp.shape[0], window):
        for j in range(0, hp.shape[1], window):
            blk = hp[i:i + window, j:j + window]
            out[i // window, j // window] = np.max(blk) if kind == "max" else np.mean(blk)
    return out


def pooled_heatmap(R, window=16, kind="max"):
    hp = np.mean(postprocess(np.asarray(R, dtype=np.float32)), axis=-1)
    return project(pool(hp, window, kind))


def overlapped_pixels(bbox, relevance, threshold):
    relevance = np.array(relevance, copy=True)
    bbox_mask = np.zeros(relevance.shape)
    bbox_mask[bbox[1]:bbox[3], bbox[0]:bbox[2]] = 1
    relevance_mask = relevance <= threshold
    if np.sum(relevance_mask > 0):
        relevance[relevance_mask] = 0
    total = np.sum(relevance)
    if total == 0:
        return 0
    ratio = 1.0 * np.sum(np.multiply(bbox_mask, relevance)) / total
    return 1. if ratio > 1 else ratio
