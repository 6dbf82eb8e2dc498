"""GPU parity: CUDA encoder relevance (through the C ABI) vs the torch-CPU oracle (oracle/encoder_ref.py)."""
import numpy as np
import pytest

from tests.util import assert_parity, rel_err, topk_cells

pytestmark = pytest.mark.gpu

RULES = {
    # name: (oracle method, oracle kwargs, analyzer name, analyzer kwargs)
    "eps": ("lrp.epsilon", dict(epsilon=0.01), "lrp.epsilon", dict(epsilon=0.01)),
    "eps_ib": ("lrp.epsilon", dict(epsilon=0.01, bias=False), "lrp.epsilon_IB", dict(epsilon=0.01)),
    "z": ("lrp.z", {}, "lrp.z", {}),
    "presetA": ("lrp.sequential_preset_a", {}, "lrp.sequential_preset_a", dict(epsilon=0.01)),
    "a1b0": ("lrp.alpha_1_beta_0", {}, "lrp.alpha_1_beta_0", {}),
    "zplus": ("lrp.z_plus", {}, "lrp.z_plus", {}),
    "zplus_fast": ("lrp.z_plus_fast", {}, "lrp.z_plus_fast", {}),
    "gradient": ("gradient", {}, "gradient", {}),
    "ixg": ("input_t_gradient", {}, "input_t_gradient", {}),
    "guided": ("guided_backprop", {}, "guided_backprop", {}),
}


def _setup(hw, n_img, img_index, seed=0):
    from lrp_imagecaptioning_b200 import synth
    from oracle import encoder_ref as ER
    W = synth.vgg16_weights(seed, bias_std=0.01)
    x = synth.images(n_img, hw, seed + 1)
    F = ER.features(x, W)
    g = np.random.default_rng(seed + 2)
    R = (F[img_index] * g.standard_normal((len(img_index),) + F.shape[1:])).astype(np.float32)
    return W, x, F, R


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("hw", [32, 64])
def test_features_match_oracle(hw, precision):
    from lrp_imagecaptioning_b200.encoder import ImageModel
    W, x, F, _ = _setup(hw, 3, [0])
    m = ImageModel(W, image_hw=hw, precision=precision)
    got = m.predict(x)
    assert rel_err(got, F) <= 1e-3 if precision == "bf16x3" else rel_err(got, F) <= 1e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("rule", sorted(RULES))
@pytest.mark.parametrize("hw", [32, 64])
def test_relevance_matches_oracle_small(hw, rule, precision):
    from lrp_imagecaptioning_b200.encoder import ImageModel
    from lrp_imagecaptioning_b200.analyzers import create_analyzer
    from oracle import encoder_ref as ER
    idx = np.array([0, 1, 1], dtype=np.int32)
    W, x, F, R = _setup(hw, 2, idx)
    om, okw, an, akw = RULES[rule]
    ref = ER.analyze(om, x[idx], R, W, **okw)
    m = ImageModel(W, image_hw=hw, precision=precision)
    got = create_analyzer(an, m, **akw).analyze_batch(x, idx, R).cpu().numpy()
    assert got.shape == ref.shape == (3, hw, hw, 3)
    for w in range(3):
        assert_parity(got[w], ref[w], "%s hw=%d %s word %d" % (rule, hw, precision, w))
        if hw >= 64:
            assert topk_cells(got[w], 5) == topk_cells(ref[w], 5)


def test_analyze_replace_mode_api():
    """analyzer.analyze([X, R]) returns an array shaped like X (innvestigate/analyzer/base.py:478-520)."""
    from lrp_imagecaptioning_b200.encoder import ImageModel
    from lrp_imagecaptioning_b200.analyzers import LRPSequentialPresetA
    from oracle import encoder_ref as ER
    idx = np.array([0, 1], dtype=np.int32)
    W, x, F, R = _setup(32, 2, idx)
    out = LRPSequentialPresetA(ImageModel(W, image_hw=32, precision="fp32"), epsilon=0.01,
                               neuron_selection_mode="replace").analyze([x, R])
    assert isinstance(out, np.ndarray) and out.shape == x.shape
    assert_parity(out, ER.analyze("lrp.sequential_preset_a", x, R, W), "analyze()")


def test_chunking_is_invisible():
    from lrp_imagecaptioning_b200.encoder import ImageModel
    from lrp_imagecaptioning_b200.analyzers import LRPEpsilon
    idx = np.array([0, 1, 1, 0, 1], dtype=np.int32)
    W, x, F, R = _setup(32, 2, idx)
    m = ImageModel(W, image_hw=32, precision="bf16x3")
    a = LRPEpsilon(m, epsilon=0.01).analyze_batch(x, idx, R).cpu().numpy()
    m.set_chunk_words(2)
    b = LRPEpsilon(m, epsilon=0.01).analyze_batch(x, idx, R).cpu().numpy()
    assert np.array_equal(a, b)


@pytest.mark.parametrize("rule,precision", [("eps", "bf16x3"), ("presetA", "bf16x3"), ("eps", "fp32")])
def test_relevance_matches_oracle_224(rule, precision):
    """BASELINE.json full image size: one image, two words."""
    from lrp_imagecaptioning_b200.encoder import ImageModel
    from lrp_imagecaptioning_b200.analyzers import create_analyzer
    from oracle import encoder_ref as ER
    idx = np.array([0, 0], dtype=np.int32)
    W, x, F, R = _setup(224, 1, idx)
    om, okw, an, akw = RULES[rule]
    ref = ER.analyze(om, x[idx], R, W, **okw)
    m = ImageModel(W, image_hw=224, precision=precision)
    got = create_analyzer(an, m, **akw).analyze_batch(x, idx, R).cpu().numpy()
    for w in range(2):
        assert_parity(got[w], ref[w], "%s 224 %s word %d" % (rule, precision, w))
        assert topk_cells(got[w], 10) == topk_cells(ref[w], 10)


def test_conservation_bias_free_224_property():
    """Size-independent property at the full size: with zero biases the z-rule conserves relevance."""
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.encoder import ImageModel
    from lrp_imagecaptioning_b200.analyzers import LRPZ
    W = [(k, np.zeros_like(b)) for k, b in synth.vgg16_weights(3)]
    x = synth.images(2, 224, 4)
    m = ImageModel(W, image_hw=224, precision="bf16x3")
    F = m.predict(x)
    idx = np.array([0, 1, 1, 0], dtype=np.int32)
    R = (F[idx] * np.random.default_rng(5).uniform(0.5, 1.5, size=(4,) + F.shape[1:])).astype(np.float32)
    out = LRPZ(m).analyze_batch(x, idx, R).cpu().numpy().astype(np.float64)
    for w in range(4):
        assert abs(out[w].sum() - float(R[w].astype(np.float64).sum())) <= 1e-3 * np.abs(R[w]).sum()
