"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *reference's own* NumPy decoder code (``/root/reference/models/explainers.py``
and ``models/model.py``) in a container that has no Keras/TensorFlow, by stubbing the
third-party imports those modules make at import time.

The reference decoder stage is plain NumPy (explainers.py:370-666, :1092-1321,
:780-832, :1452-1532), so once the imports resolve its own functions run unmodified.
This only works where ``/root/reference`` exists (the build container); it is used by
``oracle/make_golden.py`` to produce the fixtures under ``tests/golden`` and by the
CPU tests that pin ``oracle/decoder_ref.py`` against the real reference code.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LRPCAP_REFERENCE_ROOT", "/root/reference")

_STUB_TOPLEVEL = {
    "keras", "keras_applications", "tensorflow", "skimage", "matplotlib", "nltk", "tqdm",
    "psutil", "sklearn", "pycocoevalcap", "bert_score", "future", "h5py",
}

_STOP_WORDS = ["a", "an", "the", "of", "on", "in", "with", "and", "is", "are", "to", "at"]


class _StubModule(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        # must be a *class*: the reference subclasses keras.layers.Layer, Wrapper, Callback ...
        cls = type(name, (object,), {"__init__": lambda self, *a, **k: None,
                                     "__call__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _STUB_TOPLEVEL:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        if module.__name__ == "keras.backend":
            module.epsilon = lambda: 1e-7
            module.image_data_format = lambda: "channels_last"
            module.floatx = lambda: "float32"
        if module.__name__ == "nltk.corpus":
            sw = types.SimpleNamespace(words=lambda lang="english": list(_STOP_WORDS))
            module.stopwords = sw


_installed = False


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "explainers.py"))


def load_reference():
    """Returns (explainers_module, model_module) of the reference, imported under the stub."""
    global _installed
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not _installed:
        sys.meta_path.insert(0, _StubFinder())
        sys.path.insert(0, REFERENCE_ROOT)
        _installed = True
    import importlib
    E = importlib.import_module("models.explainers")
    M = importlib.import_module("models.model")
    return E, M


class _Pre:
    def __init__(self, sos, eos):
        self.SOS_TOKEN_LABEL_ENCODED = sos
        self.EOS_TOKEN_LABEL_ENCODED = eos


class _Predict:
    def __init__(self, fn):
        self.predict = fn


def make_reference_explainer(kind, variant, dec, feat, sos=1, eos=2):
    """Instantiate a reference explainer class without running its Keras constructor.

    kind: 'adaptive' | 'gridtd'; variant: 'lrp' | 'gradient'.
    dec: dict of decoder weights with our canonical names (see synth.decoder_weights).
    feat: (L, D) float32 CNN grid features handed to ``_image_model.predict``.
    """
    import numpy as np
    E, _ = load_reference()
    cls = {
        ("adaptive", "lrp"): E.ExplainImgCaptioningAdaptiveAttention,
        ("adaptive", "gradient"): E.ExplainImgCaptioningAdaptiveAttentionGradient,
        ("gridtd", "lrp"): E.ExplainImgCaptioningGridTDModel,
        ("gridtd", "gradient"): E.ExplainImgCaptioningGridTDGradient,
    }[(kind, variant)]
    o = object.__new__(cls)
    L, D = feat.shape
    side = int(np.sqrt(L))
    o.L, o.D = L, D
    o._hidden_dim = dec["hidden_dim"]
    o._embedding_dim = dec["embedding_dim"]
    o._max_caption_length = 20
    o._preprocessor = _Pre(sos, eos)
    o._image_model = _Predict(lambda x: feat.reshape(1, side, side, D))
    emb = dec["embedding"]
    embed = _Predict(lambda idx: emb[np.asarray(idx)][None])
    if kind == "adaptive":
        o._embedding = embed
        o._image_features_wieght = dec["image_features_w"]
        o._image_features_bias = dec["image_features_b"]
        o._global_img_feature_weight = dec["global_w"]
        o._global_img_feature_bias = dec["global_b"]
        o._lstm_weight_i = dec["lstm_wi"]
        o._lstm_weight_h = dec["lstm_wh"]
        o._lstm_bias = dec["lstm_b"]
        for k in ("Wv", "Wg", "V", "Wx", "Wh", "Ws"):
            setattr(o, "_" + k, dec[k])
        o._output_weight = dec["output_w"]
        o._output_bias = dec["output_b"]
    else:
        o._embedding_bm = embed
        o._image_features_weight_bm = dec["image_features_w"]
        o._image_features_bias_bm = dec["image_features_b"]
        o._global_img_feature_weight_bm = dec["global_w"]
        o._global_img_feature_bias_bm = dec["global_b"]
        o._language_lstm_weight_i = dec["lang_wi"]
        o._language_lstm_weight_h = dec["lang_wh"]
        o._language_lstm_bias = dec["lang_b"]
        o._top_down_lstm_weight_i = dec["td_wi"]
        o._top_down_lstm_weight_h = dec["td_wh"]
        o._top_down_lstm_weight_bias = dec["td_b"]
        for k in ("W_va", "W_ha", "W_a", "W_x", "W_s", "W_h"):
            setattr(o, "_" + k, dec[k])
        o._output_weight_bm = dec["output_w"]
        o._output_bias_bm = dec["output_b"]
    return o


def make_reference_lrp_inference_layer(kind, dec, vgg, mode="mean", sos=1, eos=2, word_of=None):
    """Reference LRPInferenceLayer{Adaptive,gridTD} (models/model.py:1379, :1693) without its Keras constructor: the
    NumPy decoder is the reference's own; `_image_model.predict` and `_CNN_explainer.analyze` (TensorFlow in the
    reference) are served by oracle/encoder_ref.py."""
    import numpy as np
    from oracle import encoder_ref as ER
    _, M = load_reference()
    cls = M.LRPInferenceLayerAdaptive if kind == "adaptive" else M.LRPInferenceLayergridTD
    o = object.__new__(cls)
    o._hidden_dim, o._embedding_dim = dec["hidden_dim"], dec["embedding_dim"]
    o.D = dec["D"]
    o._max_caption_length = 20
    o._preprocessor = _Pre(sos, eos)
    o._preprocessor._word_of = word_of if word_of is not None else {}
    o._EOS_ENCODED, o._SOS_ENCODER = eos, sos
    o._color_conversion = "BGRtoRGB"
    o._lrp_inference_mode = mode

    class _Img(object):
        def predict(self, x):
            f = ER.features(np.asarray(x, dtype=np.float32), vgg)
            o.L = f.shape[1] * f.shape[2]
            return f
    o._image_model = _Img()

    def _grid(x):   # grid-TD layer reads features through a Keras sub-model that already reshapes to (L, D)
        f = _Img().predict(x)
        return f.reshape(f.shape[0], -1, f.shape[-1])
    o._img_feature_input_model_bm = _Predict(_grid)
    o._CNN_explainer = type("A", (),{"analyze": staticmethod(lambda XR: ER.analyze("lrp.sequential_preset_a", XR[0], XR[1], vgg))})()
    emb = dec["embedding"]
    embed = _Predict(lambda idx: emb[np.asarray(idx)][None])
    if kind == "adaptive":
        o._embedding = embed
        o._image_features_wieght, o._image_features_bias = dec["image_features_w"], dec["image_features_b"]
        o._global_img_feature_weight, o._global_img_feature_bias = dec["global_w"], dec["global_b"]
        o._lstm_weight_i, o._lstm_weight_h, o._lstm_bias = dec["lstm_wi"], dec["lstm_wh"], dec["lstm_b"]
        for k in ("Wv", "Wg", "V", "Wx", "Wh", "Ws"):
            setattr(o, "_" + k, dec[k])
        o._output_weight, o._output_bias = dec["output_w"], dec["output_b"]
    else:
        o._embedding_bm = embed
        o._image_features_weight_bm, o._image_features_bias_bm = dec["image_features_w"], dec["image_features_b"]
        o._global_img_feature_weight_bm, o._global_img_feature_bias_bm = dec["global_w"], dec["global_b"]
        o._language_lstm_weight_i, o._language_lstm_weight_h, o._language_lstm_bias = dec["lang_wi"], dec["lang_wh"], dec["lang_b"]
        o._top_down_lstm_weight_i, o._top_down_lstm_weight_h, o._top_down_lstm_weight_bias = dec["td_wi"], dec["td_wh"], dec["td_b"]
        for k in ("W_va", "W_ha", "W_a", "W_x", "W_s", "W_h"):
            setattr(o, "_" + k, dec[k])
        o._output_weight_bm, o._output_bias_bm = dec["output_w"], dec["output_b"]
    return o
