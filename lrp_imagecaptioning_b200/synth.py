"""Seeded synthetic weights / inputs (SURVEY.md §8(d) "Synthetic inputs").

There is no network access for checkpoints or datasets, so every test and benchmark runs
on random-init weights of the reference architecture and uniform-random images:

* VGG16 conv stack to block5_conv3 (13 convs, HWIO kernels, Keras layer names
  ``block{b}_conv{c}``): He-normal ``N(0, 2/(9*Cin))``, bias ``N(0, 0.01^2)``.
* Decoder weights with the Keras tensor layouts the reference pulls out in
  ``models/explainers.py:264-278`` (adaptive) and ``:1000-1019`` (grid-TD):
  glorot-uniform kernels, orthogonal recurrent kernels, unit forget-gate bias,
  embedding ``U(-0.05, 0.05)``.
"""
import numpy as np

VGG16_CFG = [  # (keras name, Cin, Cout, pool_after)
    ("block1_conv1", 3, 64, False), ("block1_conv2", 64, 64, True),
    ("block2_conv1", 64, 128, False), ("block2_conv2", 128, 128, True),
    ("block3_conv1", 128, 256, False), ("block3_conv2", 256, 256, False), ("block3_conv3", 256, 256, True),
    ("block4_conv1", 256, 512, False), ("block4_conv2", 512, 512, False), ("block4_conv3", 512, 512, True),
    ("block5_conv1", 512, 512, False), ("block5_conv2", 512, 512, False), ("block5_conv3", 512, 512, False),
]

VGG19_CFG = [  # keras.applications.vgg19 up to block5_conv4 (models/model.py:419-421)
    ("block1_conv1", 3, 64, False), ("block1_conv2", 64, 64, True),
    ("block2_conv1", 64, 128, False), ("block2_conv2", 128, 128, True),
    ("block3_conv1", 128, 256, False), ("block3_conv2", 256, 256, False), ("block3_conv3", 256, 256, False), ("block3_conv4", 256, 256, True),
    ("block4_conv1", 256, 512, False), ("block4_conv2", 512, 512, False), ("block4_conv3", 512, 512, False), ("block4_conv4", 512, 512, True),
    ("block5_conv1", 512, 512, False), ("block5_conv2", 512, 512, False), ("block5_conv3", 512, 512, False), ("block5_conv4", 512, 512, False),
]
ENCODER_CFG = {"vgg16": VGG16_CFG, "vgg19": VGG19_CFG}

CAFFE_MEAN_BGR = np.array([103.939, 116.779, 123.68], dtype=np.float32)


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def vgg16_weights(seed=0, bias_std=0.01, cfg=None):
    """List of 13 (kernel HWIO float32, bias float32); cfg=VGG19_CFG: the 16 of VGG19."""
    g = rng(seed)
    out = []
    for _, cin, cout, _ in (cfg or VGG16_CFG):
        k = g.standard_normal((3, 3, cin, cout)).astype(np.float32) * np.float32(np.sqrt(2.0 / (9 * cin)))
        b = (g.standard_normal((cout,)) * bias_std).astype(np.float32)
        out.append((k, b))
    return out


def vgg19_weights(seed=0, bias_std=0.01):
    return vgg16_weights(seed, bias_std, VGG19_CFG)


def images(n, hw=224, seed=1):
    """Uniform [0,255] RGB images -> 'caffe' preprocess (RGB->BGR, mean subtract); NHWC float32.
    Mirrors keras vgg16.preprocess_input as used by models/preprocessors.py:43-44."""
    g = rng(seed)
    x = g.uniform(0.0, 255.0, size=(n, hw, hw, 3)).astype(np.float32)
    x = x[..., ::-1] - CAFFE_MEAN_BGR
    return np.ascontiguousarray(x, dtype=np.float32)


def _glorot(g, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return g.uniform(-lim, lim, size=shape).astype(np.float32)


def _orthogonal(g, n, m):
    a = g.standard_normal((max(n, m), max(n, m)))
    q, _ = np.linalg.qr(a)
    return np.ascontiguousarray(q[:n, :m]).astype(np.float32)


def _lstm_bias(H):
    b = np.zeros(4 * H, dtype=np.float32)
    b[H:2 * H] = 1.0  # Keras unit_forget_bias
    return b


def decoder_weights(kind, V=10000, H=512, E=512, D=512, seed=2, bias_std=0.01, out_scale=1.0):
    """Decoder weight dict. Names are ours; the mapping to the Keras tensors is in
    SURVEY.md Appendix A.1 and in ``model.py`` of this package."""
    g = rng(seed)
    d = {"kind": kind, "hidden_dim": H, "embedding_dim": E, "D": D, "vocab_size": V}
    d["image_features_w"] = _glorot(g, (D, H))
    d["image_features_b"] = (g.standard_normal(H) * bias_std).astype(np.float32)
    d["global_w"] = _glorot(g, (D, E))
    d["global_b"] = (g.standard_normal(E) * bias_std).astype(np.float32)
    d["embedding"] = g.uniform(-0.05, 0.05, size=(V, E)).astype(np.float32)
    d["output_w"] = (_glorot(g, (H, V)) * np.float32(out_scale)).astype(np.float32)
    d["output_b"] = (g.standard_normal(V) * bias_std).astype(np.float32)
    if kind == "adaptive":
        d["lstm_wi"] = _glorot(g, (2 * E, 4 * H))
        d["lstm_wh"] = np.concatenate([_orthogonal(g, H, H) for _ in range(4)], axis=1)
        d["lstm_b"] = _lstm_bias(H)
        d["Wv"] = _glorot(g, (H, H))
        d["Wg"] = _glorot(g, (H, H))
        d["Wx"] = _glorot(g, (2 * E, H))
        d["Wh"] = _glorot(g, (H, H))
        d["Ws"] = _glorot(g, (H, H))
        d["V"] = _glorot(g, (H, 1))
    elif kind == "gridtd":
        d["lang_wi"] = _glorot(g, (2 * H, 4 * H))
        d["lang_wh"] = np.concatenate([_orthogonal(g, H, H) for _ in range(4)], axis=1)
        d["lang_b"] = _lstm_bias(H)
        d["td_wi"] = _glorot(g, (H + 2 * E, 4 * H))
        d["td_wh"] = np.concatenate([_orthogonal(g, H, H) for _ in range(4)], axis=1)
        d["td_b"] = _lstm_bias(H)
        d["W_va"] = _glorot(g, (H, H))
        d["W_ha"] = _glorot(g, (H, H))
        d["W_a"] = _glorot(g, (H, 1))
        d["W_x"] = _glorot(g, (H + 2 * E, H))
        d["W_h"] = _glorot(g, (H, H))
        d["W_s"] = _glorot(g, (H, H))
    else:
        raise ValueError("kind must be 'adaptive' or 'gridtd'")
    return d


def captions(n, T, V, seed=3, sos=1, eos=2):
    """Uniform token ids (tokenizer ids, model index = id-1) in [3, V]; never SOS/EOS."""
    g = rng(seed)
    return g.integers(3, V + 1, size=(n, T)).astype(np.int32)


def features(n, L=196, D=512, seed=4, sparsity=0.5):
    """Post-ReLU-like grid features (non-negative, ~half zeros) for decoder-only tests."""
    g = rng(seed)
    f = g.standard_normal((n, L, D)).astype(np.float32)
    f = np.maximum(f - np.float32(np.quantile(f, sparsity)), 0).astype(np.float32)
    return f
