"""Native weight container standing in for the reference's Keras captioning models.

The reference explainers take a `model` object built by models/model.py (ImgCaptioningAdaptiveAttentionModel :412-604,
ImgCaptioninggridTDAdaptiveModel :606-825) and read from it: `.keras_model` (weights by Keras layer name),
`._hidden_dim`, `._embedding_dim`, `.L`, `.D`, `.img_encoder` (explainers.py:24-40).  There is no Keras here; this
class carries the same attributes and the weights under the same Keras layer/tensor names,
so a `.hdf5` -> `.npz` converter is a rename away.  Weight files are `.npz` with keys "<keras layer>/<tensor>".
"""
import numpy as np

from . import synth
from .encoder import ImageModel

ADAPTIVE_LAYER = "external_attention_rnn_wrapper_local_attention_v3_1"
GRIDTD_LAYER = "external_bottom_up_attention_adaptive_1"

# our key -> Keras "<layer>/<tensor>" name
_SHARED_NAMES = {
    "image_features_w": "image_features/kernel", "image_features_b": "image_features/bias",
    "global_w": "global_img_feature/kernel", "global_b": "global_img_feature/bias",
    "embedding": "embedding_1/embeddings", "output_w": "output/kernel", "output_b": "output/bias",
}
_ADAPTIVE_NAMES = {
    "lstm_wi": ADAPTIVE_LAYER + "/kernel", "lstm_wh": ADAPTIVE_LAYER + "/recurrent_kernel", "lstm_b": ADAPTIVE_LAYER + "/bias",
    "Wv": ADAPTIVE_LAYER + "/Wv", "Wg": ADAPTIVE_LAYER + "/Wg", "V": ADAPTIVE_LAYER + "/V",
    "Wx": ADAPTIVE_LAYER + "/Wx", "Wh": ADAPTIVE_LAYER + "/Wh", "Ws": ADAPTIVE_LAYER + "/Ws",
}
_GRIDTD_NAMES = {
    "lang_wi": GRIDTD_LAYER + "/kernel", "lang_wh": GRIDTD_LAYER + "/recurrent_kernel", "lang_b": GRIDTD_LAYER + "/bias",
    "td_wi": GRIDTD_LAYER + "/top_down_lstm_weight_i", "td_wh": GRIDTD_LAYER + "/top_down_lstm_weight_h",
    "td_b": GRIDTD_LAYER + "/top_down_lstm_weight_bias",
    "W_va": GRIDTD_LAYER + "/W_va", "W_ha": GRIDTD_LAYER + "/W_ha", "W_a": GRIDTD_LAYER + "/W_a",
    "W_x": GRIDTD_LAYER + "/W_x", "W_h": GRIDTD_LAYER + "/W_h", "W_s": GRIDTD_LAYER + "/W_s",
}


def _names(kind):
    d = dict(_SHARED_NAMES)
    d.update(_ADAPTIVE_NAMES if kind == "adaptive" else _GRIDTD_NAMES)
    return d


class CaptioningModel(object):
    """kind: 'adaptive' | 'gridtd'. vgg: list of 13 (kernel HWIO, bias). dec: decoder weight dict (synth.decoder_weights)."""

    def __init__(self, kind, vgg, dec, image_hw=224, precision="bf16x3", device="cuda:0"):
        if kind not in ("adaptive", "gridtd"):
            raise ValueError("kind must be 'adaptive' or 'gridtd'")
        if dec["kind"] != kind:
            raise ValueError("decoder weights are for %r" % dec["kind"])
        self.kind = kind
        self.dec = dec
        self.vgg = vgg
        self.image_hw = int(image_hw)
        self.precision = precision
        self.device = device
        self.img_encoder = "vgg16" if len(vgg) == 13 else "vgg19"   # models/model.py:419-421 (`img_encoder`)
        self._hidden_dim = dec["hidden_dim"]
        self._embedding_dim = dec["embedding_dim"]
        self.D = dec["D"]
        self.L = (self.image_hw // 16) ** 2
        self._vocab_size = dec["vocab_size"]
        self.image_model = ImageModel(vgg, image_hw=self.image_hw, precision=precision, device=device)

    @classmethod
    def synthetic(cls, kind, vocab_size=10000, hidden_dim=512, embedding_dim=512, image_hw=224, seed=0,
                  precision="bf16x3", device="cuda:0"):
        vgg = synth.vgg16_weights(seed)
        dec = synth.decoder_weights(kind, V=vocab_size, H=hidden_dim, E=embedding_dim, D=512, seed=seed + 1)
        return cls(kind, vgg, dec, image_hw=image_hw, precision=precision, device=device)

    # ---- weight files with Keras names
    def state_dict(self):
        out = {}
        for (k, b), name in zip(self.vgg, self.image_model.layer_names):
            out[name + "/kernel"] = k
            out[name + "/bias"] = b
        for ours, keras_name in _names(self.kind).items():
            out[keras_name] = self.dec[ours]
        return out

    def save_weights(self, path):
        np.savez(path, **self.state_dict())

    def load_weights(self, path):
        """Counterpart of keras_model.load_weights (explainers.py:27): `.npz` with Keras tensor names, or the Keras
        `.hdf5` / `.h5` file itself when h5py is importable (keras_io.read_keras_hdf5)."""
        if str(path).endswith((".hdf5", ".h5")):
            from .keras_io import read_keras_hdf5
            z = read_keras_hdf5(path)
        else:
            z = np.load(path)
        vgg = []
        for name in self.image_model.layer_names:
            vgg.append((np.asarray(z[name + "/kernel"], dtype=np.float32), np.asarray(z[name + "/bias"], dtype=np.float32)))
        dec = dict(self.dec)
        for ours, keras_name in _names(self.kind).items():
            if keras_name not in z:
                raise KeyError("weight file lacks %r" % keras_name)
            arr = np.asarray(z[keras_name], dtype=np.float32)
            if arr.shape != np.shape(self.dec[ours]):
                raise ValueError("%s: shape %s != expected %s" % (keras_name, arr.shape, np.shape(self.dec[ours])))
            dec[ours] = arr
        self.vgg, self.dec = vgg, dec
        self.image_model.close()
        self.image_model = ImageModel(vgg, image_hw=self.image_hw, precision=self.precision, device=self.device)
        return self
