"""CPU: pins the oracle.  (1) oracle/decoder_ref.py vs fixtures produced by the reference's own code
(oracle/make_golden.py -> tests/golden/decoder_*.npz); (2) where /root/reference exists, vs the reference code itself;
(3) the encoder oracle (parity unpinned at the TensorFlow boundary) vs the analyzer identities / conservation laws
iNNvestigate's own test helpers rely on (innvestigate/utils/tests/dryrun.py:187-215)."""
import os

import numpy as np
import pytest

from tests.util import linf_rel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(kind):
    z = np.load(os.path.join(GOLD, "decoder_%s.npz" % kind))
    dec = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    dec.update(kind=kind, hidden_dim=dec["image_features_w"].shape[1], embedding_dim=dec["global_w"].shape[1],
               D=dec["image_features_w"].shape[0], vocab_size=dec["embedding"].shape[0])
    return z, dec


@pytest.mark.parametrize("faithful", [False, True])
@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_decoder_oracle_matches_reference_fixture(kind, faithful):
    from oracle.decoder_ref import DecoderRef
    z, dec = _load(kind)
    cap = [int(c) for c in z["caption"]]
    o = DecoderRef(dec, faithful=faithful).forward(z["F"], cap)
    tol = 0.0 if faithful else 1e-6
    assert linf_rel(o.logits, z["logits"]) <= 1e-12
    for t in range(1, len(cap) + 1):
        r, att = o.explain(t)
        assert linf_rel(r, z["lrp_R_%d" % t]) <= tol
        assert linf_rel(att, z["lrp_att_%d" % t]) <= 1e-12
        ref_rw = z["lrp_rwords_%d" % t]
        assert o.r_words.shape == ref_rw.shape
        if ref_rw.size and np.abs(ref_rw).max() > 0:
            assert linf_rel(o.r_words, ref_rw) <= max(tol, 1e-6)
        g = o.backward(t)
        assert linf_rel(g, z["grad_R_%d" % t]) <= 1e-6
        assert linf_rel(o.r_words, z["grad_rwords_%d" % t]) <= 1e-6


def test_out_of_range_word_raises_like_reference():
    from oracle.decoder_ref import DecoderRef
    z, dec = _load("adaptive")
    o = DecoderRef(dec).forward(z["F"], [int(c) for c in z["caption"]])
    with pytest.raises(NotImplementedError):
        o.explain(len(z["caption"]) + 1)


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_decoder_oracle_matches_reference_code_when_present(kind):
    from oracle import refstub
    if not refstub.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    from lrp_imagecaptioning_b200 import synth
    from oracle.decoder_ref import DecoderRef
    dec = synth.decoder_weights(kind, V=80, H=32, E=32, D=40, seed=7)
    F = synth.features(1, L=16, D=40, seed=8)[0]
    cap = [int(c) for c in synth.captions(1, 6, 80, seed=9)[0]]
    ref = refstub.make_reference_explainer(kind, "lrp", dec, F)
    ref._forward_beam_search((None, None), cap)
    o = DecoderRef(dec, faithful=True).forward(F, cap)
    assert np.array_equal(o.logits, ref.caption_preds)
    for t in (1, 3, 6):
        r, att = ref._explain_lstm_single_word_sequence(t)
        r2, att2 = o.explain(t)
        assert np.array_equal(r, r2) and np.array_equal(att, att2)
        assert np.array_equal(np.asarray(ref.r_words), o.r_words)


# ---------------------------------------------------------------- encoder oracle: identities instead of golden vectors
def _small(bias_std, seed=0, hw=32):
    from lrp_imagecaptioning_b200 import synth
    from oracle import encoder_ref as ER
    W = synth.vgg16_weights(seed, bias_std=bias_std)
    x = synth.images(1, hw, seed + 1)
    F = ER.features(x, W)
    R = (F * np.random.default_rng(seed + 2).uniform(0.5, 1.5, size=F.shape)).astype(np.float32)
    return ER, W, x, R


@pytest.mark.parametrize("method", ["lrp.z", "lrp.epsilon", "lrp.alpha_1_beta_0", "lrp.alpha_2_beta_1", "lrp.z_plus_fast"])
def test_encoder_oracle_conserves_relevance_without_bias(method):
    ER, W, x, R = _small(0.0)
    out = ER.analyze(method, x, R, W, epsilon=1e-6).astype(np.float64)
    assert abs(out.sum() - float(R.astype(np.float64).sum())) <= 2e-3 * abs(float(R.sum()))


def test_zplus_equals_alpha1beta0_ignore_bias_and_fast_variant_on_nonnegative_input():
    """relevance_rule.py:445-455: Z+ == alpha1beta0 without bias; Z+Fast agrees when inputs are >= 0."""
    ER, W, x, R = _small(0.01)
    x = np.abs(x)
    a = ER.analyze("lrp.z_plus", x, R, W)
    b = ER.analyze("lrp.alpha_beta", x, R, W, alpha=1, beta=0, bias=False)
    c = ER.analyze("lrp.z_plus_fast", x, R, W)
    assert np.array_equal(a, b)
    assert linf_rel(c, a) <= 1e-4


def test_gradient_times_input_equals_lrp_z_without_bias():
    """relevance_analyzer.py:81-90: on bias-free ReLU nets LRP-Z == gradient*input, given consistent head seeds."""
    ER, W, x, _ = _small(0.0)
    F = ER.features(x, W)
    g = np.random.default_rng(9).standard_normal(F.shape).astype(np.float32) * (F > 0)
    a = ER.analyze("input_t_gradient", x, g, W)
    b = ER.analyze("lrp.z", x, (g * F).astype(np.float32), W)
    assert linf_rel(b, a) <= 5e-3


def test_preset_a_is_alpha1beta0_with_bias():
    ER, W, x, R = _small(0.01)
    assert np.array_equal(ER.analyze("lrp.sequential_preset_a", x, R, W), ER.analyze("lrp.alpha_1_beta_0", x, R, W))


def test_encoder_oracle_parameter_checks():
    ER, W, x, R = _small(0.01)
    with pytest.raises(ValueError):
        ER.analyze("lrp.epsilon", x, R, W, epsilon=0.0)
    with pytest.raises(ValueError):
        ER.analyze("lrp.alpha_beta", x, R, W)
    with pytest.raises(ValueError):
        ER.analyze("lrp.alpha_beta", x, R, W, alpha=2, beta=0.5)


# ---------------------------------------------------------------- LRP-inference weights (models/model.py:1641-1691, 2013-2062)
@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_lrp_inference_oracle_matches_reference_fixture(kind):
    """Fixture = output of the reference's own LRPInferenceLayer*.call under the Keras stub (encoder served by the
    torch oracle, since TensorFlow is unavailable)."""
    from oracle.make_golden import lrp_inference_case
    from oracle.lrp_inference_ref import lrp_inference_weights
    z = np.load(os.path.join(GOLD, "lrp_inference_%s.npz" % kind))
    vgg, dec, imgs, yp, word_of = lrp_inference_case(kind)
    stop = set(str(s) for s in z["stop_words"])
    for mode in ("mean", "pos_mean", "quantile"):
        got = lrp_inference_weights(dec, vgg, imgs, yp, eos=2, word_of=word_of, stop_words=stop, mode=mode)
        assert got.shape == z[mode].shape
        assert np.array_equal(got != 1, z[mode] != 1)          # same words weighted: stop word skipped, EOS stops
        assert np.abs(got - z[mode]).max() <= 1e-6
    with pytest.raises(NotImplementedError):
        lrp_inference_weights(dec, vgg, imgs, yp, eos=2, mode="bogus")


# ---------------------------------------------------------------- evaluation helpers (evaluate_bbox.py, exaimin_word.py)
def test_evaluation_oracle_matches_reference_fixture():
    """Fixture = outputs of the reference's own EvaluationBboxCOCO._get_explanation / _calculate_overlaped_pixels and
    Explainer._max_pooling / _ave_pooling (oracle/make_golden.py: main_evaluation)."""
    from oracle.make_golden import evaluation_case
    from oracle import evaluation_ref as EV
    z = np.load(os.path.join(GOLD, "evaluation.npz"))
    maps, boxes, thresholds = evaluation_case()
    for i in range(maps.shape[0]):
        assert np.abs(EV.heatmap(maps[i], "negative", True) - z["heat_negative"][i]).max() <= 1e-7
        hp = np.mean(EV.postprocess(maps[i]), axis=-1)
        assert np.abs(EV.pool(hp, 16, "max") - z["pool_max"][i]).max() <= 1e-7
        assert np.abs(EV.pool(hp, 16, "ave") - z["pool_ave"][i]).max() <= 1e-7
    for bi, (mi, x0, y0, x1, y1) in enumerate(boxes):
        for ti, th in enumerate(thresholds):
            ref = EV.overlapped_pixels([x0, y0, x1, y1], z["heat_negative"][mi], th)
            assert abs(ref - z["ratios"][bi, ti]) <= 1e-9


def test_evaluation_oracle_matches_reference_code_when_present():
    from oracle import refstub
    if not refstub.reference_available():
        pytest.skip("reference tree not present")
    from oracle.make_golden import evaluation_case, reference_evaluation_objects
    from oracle import evaluation_ref as EV
    maps, boxes, thresholds = evaluation_case()
    bb, ex = reference_evaluation_objects()
    bb._explainer.map = maps[1]
    hm, _ = bb._get_explanation((None, None), 1)
    assert np.abs(EV.heatmap(maps[1], "negative", True) - hm).max() == 0
    assert bb._calculate_overlaped_pixels([100, 3, 224, 77], hm.copy(), 0.2) == EV.overlapped_pixels([100, 3, 224, 77], hm, 0.2)
