"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the LRP-inference word weights.

Restates `LRPInferenceLayerAdaptive.call` / `LRPInferenceLayergridTD.call` (/root/reference/models/model.py:1641-1691,
:2013-2062) on top of oracle/decoder_ref.py and oracle/encoder_ref.py.  Pinned against the reference's own `call`
executed under the Keras stub with the encoder oracle plugged in as `_CNN_explainer`
(tests/test_oracle_pinning.py::test_lrp_inference_oracle_matches_reference_code, fixture tests/golden/lrp_inference_*.npz).
"""
import numpy as np

from oracle import encoder_ref as ER
from oracle.decoder_ref import DecoderRef


def _project(x):
    absmax = np.max(np.abs(x))
    if absmax == 0:
        return np.zeros(x.shape)
    return 1.0 * x / absmax


def lrp_inference_weights(dec, vgg, imgs, y_preds, eos, word_of=None, stop_words=(), mode="mean", sos=1):
    """imgs [B, hw, hw, 3]; y_preds [B, T, V] -> 1 + weights [B, T, V] (float64)."""
    if mode not in ("mean", "pos_mean", "quantile"):
        raise NotImplementedError("the lrp inference mode is not available")
    y_preds = np.asarray(y_preds)
    out = np.zeros(y_preds.shape)
    for b in range(y_preds.shape[0]):
        img = imgs[b][np.newaxis]
        cap = np.argmax(y_preds[b], axis=-1) + 1
        F = ER.features(img, vgg)[0]
        o = DecoderRef(dec, sos=sos, eos=eos).forward(F.reshape(-1, F.shape[-1]), list(cap))
        for i in range(y_preds.shape[1]):
            tok = int(cap[i])
            if word_of is not None and word_of.get(tok) in stop_words:
                continue
            if tok == eos:
                break
            rel, _ = o.explain(i + 1)
            R = ER.analyze("lrp.sequential_preset_a", img, rel, vgg)
            hp = _project(np.mean(R[..., ::-1], axis=-1)[0])
            if mode == "mean":
                s = np.mean(hp)
            elif mode == "pos_mean":
                s = np.mean(np.maximum(hp, 0))
            else:
                s = np.quantile(hp, [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9])[8]
            out[b, i, tok] = s
    return 1 + out
