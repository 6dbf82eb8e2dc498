"""GPU: LRP-inference weights and the fine-tuning step (BASELINE.json configs[4]) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Pre(object):
    SOS_TOKEN_LABEL_ENCODED = 1
    EOS_TOKEN_LABEL_ENCODED = 2

    def __init__(self, word_of):
        self._word_of = word_of


class _Provider(object):
    def __init__(self, word_of):
        self.caption_preprocessor = _Pre(word_of)


def _case(kind):
    from oracle.make_golden import lrp_inference_case
    return lrp_inference_case(kind)


@pytest.mark.parametrize("mode", ["mean", "pos_mean", "quantile"])
@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_lrp_inference_weights_match_oracle(kind, mode):
    from lrp_imagecaptioning_b200.model import CaptioningModel
    from lrp_imagecaptioning_b200 import lrp_inference as LI
    from oracle.lrp_inference_ref import lrp_inference_weights
    vgg, dec, imgs, yp, word_of = _case(kind)
    model = CaptioningModel(kind, vgg, dec, image_hw=32, precision="bf16x3")
    cls = LI.LRPInferenceLayerAdaptive if kind == "adaptive" else LI.LRPInferenceLayergridTD
    layer = cls(model, _Provider(word_of), 16, 16, 4, 512, "vgg16", mode, stop_words={"the"})
    got = layer.call([np.zeros((2, 4), np.int32), imgs, yp])
    ref = lrp_inference_weights(dec, vgg, imgs, yp, eos=2, word_of=word_of, stop_words={"the"}, mode=mode)
    assert got.shape == ref.shape == yp.shape
    assert np.array_equal(got != 1, ref != 1)
    assert np.abs(got - ref).max() <= 2e-3 * max(1e-3, np.abs(ref - 1).max()) + 1e-5
    with pytest.raises(NotImplementedError):
        cls(model, _Provider(word_of), 16, 16, 4, 512, "vgg16", "bogus")


def test_differentiable_model_matches_oracle_logits():
    """The torch training model implements the Keras step math: adaptive logits equal the oracle's forward."""
    import torch
    from lrp_imagecaptioning_b200.model import CaptioningModel
    from lrp_imagecaptioning_b200.lrp_inference import CaptionerTorch
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    vgg, dec, imgs, yp, _ = _case("adaptive")
    model = CaptioningModel("adaptive", vgg, dec, image_hw=32, precision="fp32")
    net = CaptionerTorch(model)
    cap = np.array([[5, 9, 3, 7], [4, 4, 11, 6]])
    tok_in = np.concatenate([np.ones((2, 1), int), cap[:, :-1]], axis=1)
    with torch.no_grad():
        lg = net(torch.as_tensor(tok_in).cuda(), torch.as_tensor(imgs).cuda()).cpu().numpy()
    for b in range(2):
        F = ER.features(imgs[b:b + 1], vgg)[0].reshape(-1, 512)
        o = DecoderRef(dec).forward(F, list(cap[b]))
        assert np.abs(lg[b] - o.logits).max() <= 1e-3 * np.abs(o.logits).max()


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_fine_tune_step_runs_and_learns(kind):
    from lrp_imagecaptioning_b200.model import CaptioningModel
    from lrp_imagecaptioning_b200.lrp_inference import LRPInferenceTrainer
    vgg, dec, imgs, yp, word_of = _case(kind)
    model = CaptioningModel(kind, vgg, dec, image_hw=32, precision="bf16x3")
    tr = LRPInferenceTrainer(model, _Provider(word_of), "mean", learning_rate=1e-4)
    g = np.random.default_rng(0)
    cap = g.integers(3, 29, size=(2, 4))
    tok_in = np.concatenate([np.ones((2, 1), int), cap[:, :-1]], axis=1)
    y = np.eye(30, dtype=np.float32)[cap - 1]
    losses = [tr.step(tok_in, imgs, y) for _ in range(6)]
    assert all(np.isfinite(losses)) and min(losses[1:]) < losses[0], losses
    assert tr.explained_words_total > 0


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_on_device_path_equals_the_layer_call(kind):
    """The fine-tune step's on-device pieces against the mirrored reference interface: (1) the engine's prediction
    (teacher-forced, arg-max mode) = argmax of the torch model's logits + 1; (2) the sparse weights = the non-trivial
    entries of `LRPInferenceLayer*.call`; (3) handles updated device-to-device explain like freshly created ones."""
    import torch
    from lrp_imagecaptioning_b200.model import CaptioningModel
    from lrp_imagecaptioning_b200 import lrp_inference as LI
    vgg, dec, imgs, yp, word_of = _case(kind)
    model = CaptioningModel(kind, vgg, dec, image_hw=32, precision="bf16x3")
    tr = LI.LRPInferenceTrainer(model, _Provider(word_of), "mean", learning_rate=1e-3, stop_words={"the"})
    g = np.random.default_rng(1)
    cap = g.integers(3, 29, size=(2, 4))
    tok_in = np.concatenate([np.ones((2, 1), int), cap[:, :-1]], axis=1)
    with torch.no_grad():
        lg = tr.net(torch.as_tensor(tok_in).cuda(), torch.as_tensor(imgs).cuda())
    want = (lg.argmax(dim=-1) + 1).cpu().numpy()
    tr.net.sync_engine(tr.layer._engine)
    pred = tr.predict(tok_in, torch.as_tensor(imgs).cuda())
    assert np.array_equal(pred, want)
    b, t, col, sc = tr.layer.sparse_weights(torch.as_tensor(imgs).cuda(), pred, features_ready=True)
    y_pred = np.zeros((2, 4, 30), dtype=np.float32)
    np.put_along_axis(y_pred, (pred - 1)[..., None], 1.0, axis=-1)
    dense = tr.layer.call([tok_in, imgs, y_pred])
    sparse = np.ones_like(dense)
    keep = col < 30
    sparse[b[keep], t[keep], col[keep]] += sc[keep]
    assert np.array_equal(dense != 1, sparse != 1)
    assert np.abs(dense - sparse).max() <= 1e-6
    # one optimizer step, then: device-updated handles == handles created from the exported weights
    y = np.eye(30, dtype=np.float32)[cap - 1]
    tr.step(tok_in, imgs, y)
    tr.net.sync_engine(tr.layer._engine)
    pred2 = tr.predict(tok_in, torch.as_tensor(imgs).cuda())
    s_dev = tr.layer.sparse_weights(torch.as_tensor(imgs).cuda(), pred2, features_ready=True)
    fresh_model = tr.net.export(model)
    cls = LI.LRPInferenceLayerAdaptive if kind == "adaptive" else LI.LRPInferenceLayergridTD
    fresh = cls(fresh_model, _Provider(word_of), 16, 16, 4, 512, "vgg16", "mean", stop_words={"the"}, overflow="skip")
    s_new = fresh.sparse_weights(torch.as_tensor(imgs).cuda(), pred2)
    for a, c in zip(s_dev, s_new):
        assert np.array_equal(a, c)
