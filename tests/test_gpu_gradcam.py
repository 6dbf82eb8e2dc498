"""GPU parity: Grad-CAM kernels vs the scipy restatement of skimage.pyramid_expand (parity unpinned, see oracle)."""
import numpy as np
import pytest

from tests.util import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fh", [4, 14])
def test_gradcam_matches_oracle(fh):
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.gradcam import grad_cam_batch, grad_cam, scale_maps
    from oracle.gradcam_ref import grad_cam as ref_cam
    import torch
    L, D, N = fh * fh, 512, 2
    F = synth.features(N, L=L, D=D, seed=3)
    idx = np.array([1, 0, 1], dtype=np.int32)
    g = np.random.default_rng(4).standard_normal((3, L, D)).astype(np.float32)
    cam = grad_cam_batch(F, idx, g).cpu().numpy()
    assert cam.shape == (3, fh * 16, fh * 16)
    for w in range(3):
        assert_parity(cam[w], ref_cam(F[idx[w]], g[w], L, D), "gradcam fh=%d word %d" % (fh, w), rel_tol=1e-4, sum_tol=None)
    one = grad_cam(F[1], g[0].reshape(fh, fh, D), L, D)
    assert np.array_equal(one, cam[0])
    maps = torch.ones((3, fh * 16, fh * 16, 3), device="cuda") * 2.0
    scale_maps(maps, torch.from_numpy(cam).cuda())
    assert np.allclose(maps.cpu().numpy(), 2.0 * cam[..., None])
