"""Unit parity of the convolution kernels through the C ABI (lrpcap_debug_conv) against a float64 numpy contraction.

Covers both tcgen05 kernels -- the generic one (csrc/tc_conv.cu) and the vertical-halo variant for wide shallow
layers (csrc/tc_conv_vh.cu: W % 16 == 0, H % 16 == 0, 64 / 128 output channels) -- plus the fp32 SIMT kernel.
The contraction is the one iNNvestigate emits for a conv layer and its GradientWRT (innvestigate/layers.py:138-157)."""
import numpy as np
import pytest

from lrp_imagecaptioning_b200 import _lib

pytestmark = pytest.mark.gpu


def ref_conv(A, B, taps):
    A = A.astype(np.float64)
    B = B.astype(np.float64)
    items, H, W, C = A.shape
    if taps == 1:
        return A @ B[0]
    out = np.zeros((items, H, W, B.shape[-1]))
    Ap = np.pad(A, ((0, 0), (1, 1), (1, 1), (0, 0)))
    for t in range(9):
        dy, dx = t // 3, t % 3
        out += Ap[:, dy:dy + H, dx:dx + W, :] @ B[t]
    return out


# (items, H, W, C, Nout, taps)
GENERIC = [(1, 8, 16, 64, 64, 1), (1, 14, 14, 128, 256, 9), (3, 28, 28, 256, 256, 9), (1, 56, 56, 128, 64, 9),
           (1, 4, 4, 64, 128, 9), (1, 2, 2, 512, 512, 9), (5, 7, 7, 64, 64, 1), (1, 8, 16, 64, 64, 9)]
VERTICAL_HALO = [(2, 16, 16, 64, 64, 9), (1, 16, 32, 128, 128, 9), (2, 32, 32, 64, 64, 9), (1, 32, 48, 128, 64, 9),
                 (3, 48, 32, 256, 128, 9), (150, 16, 16, 64, 64, 9), (1, 112, 112, 128, 64, 9)]
# N = 256 shapes whose 128-pixel tiles pair up: the CTA-pair (cta_group::2) kernels; (40, 14, 14, ..) gives 80 pair tiles,
# more than the 74 pairs of a B200, so some pairs run two tiles through both accumulator buffers
PAIRS = [(4, 14, 14, 512, 512, 9), (2, 28, 28, 512, 256, 9), (40, 14, 14, 256, 512, 9), (2, 14, 14, 64, 256, 1), (1, 56, 56, 256, 256, 9),
         (4, 14, 14, 256, 128, 9),                                   # N = 128 generic kernel
         # vertical-halo kernel, 16 x 16 pixel tiles in pairs (150 items: more pair tiles than CTA pairs)
         (2, 16, 16, 64, 64, 9), (1, 16, 32, 128, 128, 9), (2, 32, 32, 64, 64, 9), (1, 32, 48, 128, 64, 9), (3, 48, 32, 256, 128, 9),
         (150, 16, 16, 64, 64, 9), (2, 112, 112, 128, 64, 9)]
TOL = {_lib.PREC_FP32_SIMT: 2e-6, _lib.PREC_BF16X3_TC: 3e-5, 2: 2e-6, 3: 2e-6, 4: 3e-5, 5: 1e-4}


@pytest.mark.parametrize("prec", [_lib.PREC_FP32_SIMT, _lib.PREC_BF16X3_TC, 2, 3, 4, 5], ids=["simt", "tc", "tc3", "f16x2", "h1x2", "h1f8"])
@pytest.mark.parametrize("shape", GENERIC + VERTICAL_HALO, ids=lambda s: "x".join(map(str, s)))
def test_conv_matches_float64(prec, shape):
    items, H, W, C, Nout, taps = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 31))
    A = rng.standard_normal((items, H, W, C)).astype(np.float32)
    B = (rng.standard_normal((taps, C, Nout)) / np.sqrt(taps * C)).astype(np.float32)
    got = _lib.debug_conv(prec, A, B, taps)
    if prec == 4:   # two-product mode: the A operand is one fp16 plane by construction; the kernel must be exact beyond that
        A = A.astype(np.float16).astype(np.float32)
    ref = ref_conv(A, B, taps)
    assert np.isfinite(got).all()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < TOL[prec], (shape, err)


def test_tap_geometry_vertical_halo():
    """One-hot taps with identity weights: the output must be the input shifted by the tap (padding = zeros)."""
    H, W, C = 32, 32, 64
    A = (np.arange(H * W)[:, None] + np.arange(C)[None, :] / 128.0).reshape(1, H, W, C).astype(np.float32)
    for t in range(9):
        B = np.zeros((9, C, C), dtype=np.float32)
        B[t] = np.eye(C)
        got = _lib.debug_conv(_lib.PREC_BF16X3_TC, A, B, 9)
        ref = ref_conv(A, B, 9)
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5, t


@pytest.mark.parametrize("promote", [0, 9], ids=["plain", "promoted"])
@pytest.mark.parametrize("prec", [_lib.PREC_BF16X3_TC, 4, 5], ids=["tc", "h1x2", "h1f8"])
@pytest.mark.parametrize("shape", PAIRS, ids=lambda s: "x".join(map(str, s)))
def test_cta_pair_conv_matches_the_one_cta_kernel(prec, shape, promote, monkeypatch):
    """The cta_group::2 kernels (two CTAs share one M = 256 MMA, each staging half of the weight rows) against float64 and,
    bit for bit, against the one-CTA kernel (LRPCAP_TC_2SM=0 is read once per process, so that arm runs in a child)."""
    import subprocess, sys, os, tempfile
    items, H, W, C, Nout, taps = shape
    rng = np.random.default_rng(7)
    A = rng.standard_normal((items, H, W, C)).astype(np.float32)
    B = (rng.standard_normal((taps, C, Nout)) / np.sqrt(taps * C)).astype(np.float32)
    if promote:
        monkeypatch.setenv("LRPCAP_DEBUG_CONV_PROMOTE", str(promote))
    got = _lib.debug_conv(prec, A, B, taps)
    Ar = A.astype(np.float16).astype(np.float32) if prec == 4 else A
    ref = ref_conv(Ar, B, taps).reshape(got.shape)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert np.isfinite(got).all() and err <= TOL[prec], err
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "A.npy"), A); np.save(os.path.join(d, "B.npy"), B)
        code = ("import numpy as np, sys; sys.path.insert(0, %r); from lrp_imagecaptioning_b200 import _lib; "
                "np.save(%r, _lib.debug_conv(%d, np.load(%r), np.load(%r), %d))"
                % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(d, "o.npy"), prec,
                   os.path.join(d, "A.npy"), os.path.join(d, "B.npy"), taps))
        env = dict(os.environ, LRPCAP_TC_2SM="0")
        subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
        one = np.load(os.path.join(d, "o.npy"))
    if prec == 4 and Nout == 128 and (H % 16 or promote):
        # N = 128 two-product launch of the generic kernel: one CTA keeps A*hi and A*lo in separate accumulator columns and adds
        # them in the epilogue, the pair accumulates both into one -- the same products in a different fp32 summation order
        assert np.abs(one - got).max() <= 1e-5 * np.abs(one).max()   # tensor-core accumulation rounds toward zero per step
    else:
        assert np.array_equal(one, got), np.abs(one - got).max()
