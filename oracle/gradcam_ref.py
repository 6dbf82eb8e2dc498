"""TEST INFRASTRUCTURE ONLY -- CPU oracle for Grad-CAM (reference models/explainers.py:939-949).

PARITY UNPINNED: the reference calls scikit-image's `pyramid_expand(cam, upscale=16, sigma=20, multichannel=False)`,
which is neither vendored nor installable here (version unpinned, README.md:9-11).  Its published algorithm
(skimage/transform/pyramids.py: `resize(order=1, mode='reflect', anti_aliasing=False)` then
`scipy.ndimage.gaussian_filter(sigma, mode='reflect')`; skimage 'reflect' maps to ndimage 'mirror' inside warp) is
restated with scipy.ndimage."""
import numpy as np
from scipy import ndimage as ndi


def pyramid_expand(image, upscale=16, sigma=20.0):
    image = np.asarray(image, dtype=np.float64)
    h, w = image.shape
    oy = (np.arange(h * upscale) + 0.5) / upscale - 0.5
    ox = (np.arange(w * upscale) + 0.5) / upscale - 0.5
    yy, xx = np.meshgrid(oy, ox, indexing="ij")
    resized = ndi.map_coordinates(image, [yy, xx], order=1, mode="mirror")
    return ndi.gaussian_filter(resized, sigma, mode="reflect", truncate=4.0)


def grad_cam(img_feature, grads, L, D):
    side = int(np.sqrt(L))
    weights = np.mean(np.asarray(grads).reshape(side, side, D), axis=(0, 1))
    conv_output = np.asarray(img_feature).reshape(side, side, D)
    cam = np.zeros((side, side), dtype=np.float32)
    for i, w in enumerate(weights):
        cam += w * conv_output[:, :, i]
    cam = pyramid_expand(cam, upscale=16, sigma=20)
    cam = np.maximum(cam, 0)
    return cam / (np.max(np.abs(cam)) + 1e-6)
