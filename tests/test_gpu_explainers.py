"""GPU: the drop-in explainer classes (mirror of models/explainers.py) end to end against the oracle pipeline
(reference call sequence: explain_image.py:45-87)."""
import numpy as np
import pytest

from tests.util import assert_parity, linf_rel

pytestmark = pytest.mark.gpu

HW, V, H = 64, 120, 64


class _Pre(object):
    SOS_TOKEN_LABEL_ENCODED = 1
    EOS_TOKEN_LABEL_ENCODED = 2


class _Provider(object):
    caption_preprocessor = _Pre()


def _model(kind, seed=0):
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.model import CaptioningModel
    vgg = synth.vgg16_weights(seed)
    dec = synth.decoder_weights(kind, V=V, H=H, E=H, D=512, seed=seed + 1)
    return CaptioningModel(kind, vgg, dec, image_hw=HW, precision="bf16x3"), vgg, dec


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_lrp_explainer_matches_oracle_pipeline(kind, tmp_path):
    from lrp_imagecaptioning_b200 import synth, explainers as E
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    model, vgg, dec = _model(kind)
    path = str(tmp_path / "w.npz")
    model.save_weights(path)
    cls = E.ExplainImgCaptioningAdaptiveAttention if kind == "adaptive" else E.ExplainImgCaptioningGridTDModel
    ex = cls(model, path, _Provider(), 20)
    img = synth.images(1, HW, 5)
    cap = [int(c) for c in synth.captions(1, 6, V, seed=6)[0]]
    ex._forward_beam_search((None, img), cap)
    rel, att = ex._explain_sentence()
    assert len(rel) == len(cap) - 1 and att.shape == (len(cap) - 1, ex.L)
    F = ER.features(img, vgg)[0].reshape(-1, 512)
    o = DecoderRef(dec).forward(F, cap)
    side = HW // 16
    errs = []
    for i, r in enumerate(rel):
        assert r.shape == (1, side, side, 512) and r.dtype == np.float32
        rF, a = o.explain(i + 1)
        assert_parity(r, rF, "%s explainer R_F word %d" % (kind, i + 1), rel_tol=2e-3, sum_tol=None)
        assert_parity(att[i], o.attention[1:-1][i], "%s explainer attention %d" % (kind, i + 1), rel_tol=1e-4, sum_tol=None)
        heat = ex._explain_CNN(img, r)
        assert heat.shape == (1, HW, HW, 3)
        errs.append(linf_rel(heat, ER.analyze("lrp.sequential_preset_a", img, rF, vgg)))
    assert np.median(errs) <= 2e-3 and max(errs) <= 8e-2, errs
    r1, a1 = ex._explain_lstm_single_word_sequence(3)
    assert np.array_equal(r1, rel[2])
    rF, _ = o.explain(3)
    if len(o.r_words):
        assert_parity(ex.r_words, o.r_words, "%s r_words" % kind, sum_tol=None)
    with pytest.raises(NotImplementedError):
        ex._explain_lstm_single_word_sequence(len(cap) + 1)
    batch = ex.explain_sentence_to_pixels()
    assert batch.shape == (len(cap) - 1, HW, HW, 3)
    assert np.array_equal(batch[0], ex._explain_CNN(img, rel[0])[0])


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_gradient_family_explainers(kind):
    from lrp_imagecaptioning_b200 import synth, explainers as E
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    from oracle.gradcam_ref import grad_cam as ref_cam
    model, vgg, dec = _model(kind, seed=3)
    names = {"adaptive": ("ExplainImgCaptioningAdaptiveAttentionGradient", "ExplainImgCaptioningAdaptiveAttentionInputTimesGradient",
                          "ExplainImgCaptioningAdaptiveAttentionGuidedGradcam"),
             "gridtd": ("ExplainImgCaptioningGridTDGradient", "ExplainImgCaptioningGridTDGradientTimesInput",
                        "ExplainImgCaptioningGridTDGuidedGradcam")}[kind]
    img = synth.images(1, HW, 7)
    cap = [int(c) for c in synth.captions(1, 5, V, seed=8)[0]]
    F = ER.features(img, vgg)[0].reshape(-1, 512)
    o = DecoderRef(dec).forward(F, cap)
    for name, method in zip(names, ("gradient", "input_t_gradient", "guided_backprop")):
        ex = getattr(E, name)(model, None, _Provider(), 20)
        ex._forward_beam_search((None, img), cap)
        rel = ex._explain_sentence()
        assert len(rel) == len(cap) - 1
        g = o.backward(2)
        assert_parity(rel[1], g, "%s decoder gradient" % name, rel_tol=2e-3, sum_tol=None)
        assert np.array_equal(ex._lstm_decoder_backward(2), rel[1])
        heat = ex._explain_CNN(img, rel[1])
        ref = ER.analyze(method, img, g, vgg)
        if method == "guided_backprop":
            ref = (ref[0] * ref_cam(F, g[0], F.shape[0], 512)[..., None])[None]
        assert heat.shape == ref.shape
        assert linf_rel(heat, ref) <= 8e-2


def test_beam_search_matches_host_reimplementation():
    """explainers.py:51-120: same search on the oracle's logits (Keras logits = explainer logits for adaptive)."""
    import heapq
    from lrp_imagecaptioning_b200 import synth, explainers as E
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    model, vgg, dec = _model("adaptive", seed=5)
    ex = E.ExplainImgCaptioningAdaptiveAttention(model, None, _Provider(), 6)
    img = synth.images(1, HW, 9)
    beam = 3
    got = ex._beam_search((None, img), beam)
    assert len(got) == beam and all(c[-1] == 2 or len(c) == 7 for c in got)
    F = ER.features(img, vgg)[0].reshape(-1, 512)

    def logp_next(prefix):
        o = DecoderRef(dec).forward(F, prefix + [1])
        row = o.logits[-1] - o.logits[-1].max()
        return row - np.log(np.exp(row).sum())
    partial, complete = [(0.0, [1, 2])], []

    def push(h, it):
        heapq.heappush(h, it) if len(h) < beam else heapq.heappushpop(h, it)
    for _ in range(6):
        prev, partial = sorted(partial, reverse=True), []
        for lp_prev, sent in prev:
            row = logp_next(sent[1:-1])
            for w in np.argsort(row)[-beam:]:
                lp = float(row[w] + lp_prev)
                push(partial, (lp, sent[:-1] + [int(w) + 1, sent[-1]]))
                if int(w) + 1 == 2:
                    push(complete, (lp, sent))
    top_p, top_c = sorted(partial, reverse=True), sorted(complete, reverse=True)
    want = [(top_c[r] if r < len(top_c) else top_p[r])[1][1:] for r in range(beam)]
    assert got == want


def test_greedy_is_beam_one():
    from lrp_imagecaptioning_b200 import synth, explainers as E
    model, vgg, dec = _model("gridtd", seed=6)
    ex = E.ExplainImgCaptioningGridTDModel(model, None, _Provider(), 5)
    caps = ex._beam_search((None, synth.images(2, HW, 9)), 1)
    assert len(caps) == 1 and len(caps[0]) == 2 and all(len(c) >= 1 for c in caps[0])


def test_streamed_engine_lanes_give_identical_maps():
    """Independent lanes (own handles, stream and host thread per block of images) must reproduce the single-engine
    result bit for bit, resident and through the host-buffer C-ABI call."""
    import torch
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.engine import ExplainEngine, StreamedEngine
    model, _, _ = _model("gridtd", seed=3)
    x = synth.images(5, HW, 21)
    T = 4
    one = ExplainEngine(model)
    ref_maps, ref_cap = one.explain_batch(torch.from_numpy(x).cuda(), T=T, greedy=True)
    ref_maps = ref_maps.cpu().numpy()
    se = StreamedEngine(model, lanes=3, chunk_words=7)
    maps, cap = se.explain_batch(torch.from_numpy(x).cuda(), T, greedy=True)
    torch.cuda.synchronize()
    got = torch.cat(maps, 0).cpu().numpy()
    assert np.array_equal(cap, ref_cap)
    assert np.array_equal(got, ref_maps)
    cap_h = np.zeros((5, T), dtype=np.int32)
    host = se.explain_batch_host(x, cap_h, greedy=True)
    assert np.array_equal(cap_h, ref_cap)
    assert np.array_equal(host, ref_maps)


def test_host_buffer_call_equals_resident_path_with_shrinking_tail_chunks():
    """lrpcap_explain_batch_host streams every chunk of maps to the host and halves the last chunks (>= 128 words):
    the partition must still cover every word exactly once, bit-identically to the device-resident path."""
    import torch
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.engine import ExplainEngine
    from lrp_imagecaptioning_b200.model import CaptioningModel
    vgg = synth.vgg16_weights(2)
    dec = synth.decoder_weights("adaptive", V=V, H=H, E=H, D=512, seed=5)
    model = CaptioningModel("adaptive", vgg, dec, image_hw=32, precision="bf16x3")
    model.image_model.set_chunk_words(192)
    eng = ExplainEngine(model)
    x = synth.images(18, 32, 4)
    T = 20                                           # 360 words: chunks 192, then 96 + 64 + ... of the 168-word remainder
    maps, cap = eng.explain_batch(torch.from_numpy(x).cuda(), T=T, greedy=True)
    cap_h = np.zeros((18, T), dtype=np.int32)
    host = eng.explain_batch_host(x, cap_h, greedy=True)
    assert np.array_equal(cap_h, cap)
    assert np.array_equal(host, maps.cpu().numpy())
