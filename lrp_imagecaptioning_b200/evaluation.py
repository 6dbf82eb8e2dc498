"""Explanation-evaluation reductions on the device (include/lrpcap.h "evaluation reductions").

Host mirror of the helpers the reference applies to every pixel map before evaluating it:
  exaimin_word.py:64-77   Explainer._max_pooling / _ave_pooling
  exaimin_word.py:80-128  project / heat map of a single word
  evaluate_bbox.py:59-86  EvaluationBboxCOCO._get_explanation (negative part, relu, channel mean, project)
  evaluate_bbox.py:191-208 EvaluationBboxCOCO._calculate_overlaped_pixels
Inputs are the [W, hw, hw, 3] pixel maps the encoder produced (torch cuda tensors stay on the device; numpy is copied)."""
import numpy as np
import torch

from . import _lib

MODES = {"mean": 0, "negative": 1, "positive": 2}
THRESHOLDS = (0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)   # evaluate_bbox.py:251


def _cuda(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return x.to(torch.device(device)).contiguous().float()


def heatmaps(R_pix, mode="mean", shift_negative=False, window=1, pooling="max", want_means=False, device="cuda:0"):
    """R_pix [W, hw, hw, 3] -> projected heat maps [W, hw/window, hw/window] (torch cuda) [, their means (numpy [W])].

    mode 'mean' is exaimin_word's heat map, 'negative' evaluate_bbox's (pass shift_negative=True for its `project`);
    window=16 with pooling 'max' | 'ave' reproduces `_explain_single_word_pooling`."""
    if mode not in MODES:
        raise ValueError("mode must be one of %s" % sorted(MODES))
    if pooling not in ("max", "ave"):
        raise ValueError("pooling must be 'max' or 'ave'")
    m = _cuda(R_pix, device)
    if m.dim() != 4 or m.shape[1] != m.shape[2] or m.shape[3] != 3:
        raise ValueError("expected pixel maps [W, hw, hw, 3]")
    W, hw = m.shape[0], m.shape[1]
    if window < 1 or hw % window:
        raise ValueError("window must divide the map size")
    out = torch.empty((W, hw // window, hw // window), dtype=torch.float32, device=m.device)
    means = np.zeros(W, dtype=np.float32) if want_means else None
    stream = _lib.c_void_p(torch.cuda.current_stream(m.device).cuda_stream)
    _lib.check(_lib.load().lrpcap_heatmaps(_lib.c_void_p(m.data_ptr()), W, hw, MODES[mode], int(bool(shift_negative)),
                                           int(window), 0 if pooling == "max" else 1, _lib.c_void_p(out.data_ptr()),
                                           _lib.fptr(means) if want_means else None, stream))
    return (out, means) if want_means else out


def bbox_correctness(heat, boxes, thresholds=THRESHOLDS, device="cuda:0"):
    """heat [M, hw, hw]; boxes: rows (map index, x0, y0, x1, y1) -> ratios [n_boxes, n_thresholds] (numpy float32).
    ratio = relevance mass inside the box / total mass over the pixels above the threshold (0 if none, capped at 1)."""
    h = _cuda(heat, device)
    if h.dim() != 3 or h.shape[1] != h.shape[2]:
        raise ValueError("expected heat maps [M, hw, hw]")
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int32).reshape(-1, 5))
    th = np.ascontiguousarray(np.asarray(thresholds, dtype=np.float32).reshape(-1))
    out = np.zeros((b.shape[0], th.shape[0]), dtype=np.float32)
    stream = _lib.c_void_p(torch.cuda.current_stream(h.device).cuda_stream)
    _lib.check(_lib.load().lrpcap_bbox_correctness(_lib.c_void_p(h.data_ptr()), h.shape[0], h.shape[1], _lib.iptr(b), b.shape[0],
                                                   _lib.fptr(th), th.shape[0], _lib.fptr(out), stream))
    return out


def category_means(R_pix, device="cuda:0"):
    """np.mean of every projected 'mean' heat map (exaimin_word.py:446: `np.mean(hp)` per (image, category word))."""
    return heatmaps(R_pix, "mean", want_means=True, device=device)[1]
