"""Grad-CAM on the device (include/lrpcap.h "Grad-CAM" section; reference models/explainers.py:930-949)."""
import numpy as np
import torch

from . import _lib


def grad_cam_batch(features, img_index, grads, upscale=16, sigma=20.0, device="cuda:0"):
    """features [N, L, D], grads [W, L, D] (torch cuda or numpy) -> cam [W, fh*upscale, fh*upscale] torch cuda."""
    dev = torch.device(device)

    def cu(x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return x.to(dev).contiguous().float()
    f, g = cu(features), cu(grads)
    f = f.reshape(f.shape[0], -1, f.shape[-1])
    g = g.reshape(g.shape[0], -1, g.shape[-1])
    L, D = f.shape[1], f.shape[2]
    fh = int(round(np.sqrt(L)))
    if fh * fh != L or g.shape[1] != L or g.shape[2] != D:
        raise ValueError("features / grads must share a square grid and depth")
    idx = np.ascontiguousarray(np.asarray(img_index), dtype=np.int32)
    out = torch.empty((g.shape[0], fh * upscale, fh * upscale), dtype=torch.float32, device=dev)
    stream = _lib.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.load().lrpcap_gradcam(_lib.c_void_p(f.data_ptr()), _lib.iptr(idx), _lib.c_void_p(g.data_ptr()),
                                          g.shape[0], fh, D, int(upscale), float(sigma), _lib.c_void_p(out.data_ptr()), stream))
    return out


def grad_cam(img_feature, grads, L, D, upscale=16, sigma=20.0):
    """Single-word form with the reference's signature: img_feature (L, D), grads (h, w, D) -> (h*16, w*16) numpy."""
    cam = grad_cam_batch(np.asarray(img_feature, dtype=np.float32).reshape(1, L, D), [0],
                         np.asarray(grads, dtype=np.float32).reshape(1, L, D), upscale, sigma)
    return cam[0].cpu().numpy()


def scale_maps(maps, cam):
    """maps [W, hw, hw, 3] (cuda, modified in place) *= cam [W, hw, hw]."""
    stream = _lib.c_void_p(torch.cuda.current_stream(maps.device).cuda_stream)
    _lib.check(_lib.load().lrpcap_scale_maps(_lib.c_void_p(maps.data_ptr()), _lib.c_void_p(cam.data_ptr()), maps.shape[0],
                                             maps.shape[1], stream))
    return maps
