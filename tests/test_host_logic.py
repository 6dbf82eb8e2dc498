"""CPU: host-side mirror of the reference interfaces (parameter validation, registries, sharding, weight files)."""
import numpy as np
import pytest
import torch


def _image_model():
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.encoder import ImageModel
    return ImageModel(synth.vgg16_weights(0), image_hw=32, precision="fp32")


def test_analyzer_registry_and_parameter_errors():
    """Same names and ValueErrors as innvestigate (analyzer/__init__.py:35-99, relevance_based/utils.py:52-129)."""
    from lrp_imagecaptioning_b200 import analyzers as A
    m = _image_model()
    for name in ("lrp.epsilon", "lrp.z", "lrp.alpha_beta", "lrp.alpha_1_beta_0", "lrp.alpha_2_beta_1", "lrp.z_plus",
                 "lrp.z_plus_fast", "lrp.sequential_preset_a", "gradient", "input_t_gradient", "guided_backprop"):
        assert name in A.analyzers
    with pytest.raises(ValueError):
        A.create_analyzer("lrp.epsilon", m, epsilon=0)
    with pytest.raises(ValueError):
        A.LRPAlphaBeta(m)
    with pytest.raises(ValueError):
        A.LRPAlphaBeta(m, alpha=0.5)
    with pytest.raises(ValueError):
        A.LRPAlphaBeta(m, alpha=2, beta=0.5)
    with pytest.raises(ValueError):
        A.Gradient(m, neuron_selection_mode="bogus")
    with pytest.raises(ValueError):
        A.Gradient(m, postprocess="cube")
    a = A.LRPAlphaBeta(m, beta=1)
    assert (a._alpha, a._beta) == (2, 1)
    r = A.LRPSequentialPresetA(m, epsilon=0.01, neuron_selection_mode="replace")._rule()
    assert (r.alpha, r.beta, r.bias) == (1.0, 0.0, True)
    with pytest.raises(A.NotAnalyzeableModelException):
        A.LRPEpsilon(object())


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from lrp_imagecaptioning_b200 import _lib, synth
    from lrp_imagecaptioning_b200.analyzers import LRPEpsilon
    m = _image_model()
    with pytest.raises(_lib.LrpcapError):
        LRPEpsilon(m, epsilon=0.01).analyze([synth.images(1, 32), np.zeros((1, 2, 2, 512), np.float32)])


def test_image_model_shape_checks():
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.encoder import ImageModel
    w = synth.vgg16_weights(0)
    with pytest.raises(ValueError):
        ImageModel(w[:12])
    bad = list(w)
    bad[3] = (bad[3][0][..., :5], bad[3][1])
    with pytest.raises(ValueError):
        ImageModel(bad)


def test_shard_images_partitions_everything_once():
    from lrp_imagecaptioning_b200.engine import shard_images, word_list
    for n in (0, 1, 7, 64, 513):
        for ws in (1, 2, 3, 8):
            parts = [shard_images(n, r, ws) for r in range(ws)]
            assert np.array_equal(np.concatenate(parts), np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    wi, wt = word_list(3, 4, lengths=[2, 0, 4])
    assert list(wi) == [0, 0, 2, 2, 2, 2] and list(wt) == [1, 2, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        shard_images(4, 2, 2)


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
def test_weight_file_roundtrip_uses_keras_names(kind, tmp_path):
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.model import CaptioningModel, ADAPTIVE_LAYER, GRIDTD_LAYER
    vgg = synth.vgg16_weights(0)
    dec = synth.decoder_weights(kind, V=50, H=16, E=16, D=512, seed=1)
    m = CaptioningModel(kind, vgg, dec, image_hw=32, precision="fp32")
    sd = m.state_dict()
    assert "block5_conv3/kernel" in sd and "image_features/kernel" in sd and "embedding_1/embeddings" in sd
    assert ((ADAPTIVE_LAYER if kind == "adaptive" else GRIDTD_LAYER) + "/recurrent_kernel") in sd
    p = str(tmp_path / "w.npz")
    m.save_weights(p)
    dec2 = synth.decoder_weights(kind, V=50, H=16, E=16, D=512, seed=99)
    m2 = CaptioningModel(kind, synth.vgg16_weights(5), dec2, image_hw=32, precision="fp32").load_weights(p)
    for k, v in m.state_dict().items():
        assert np.array_equal(m2.state_dict()[k], v)
    assert (m2.L, m2.D, m2._hidden_dim, m2._embedding_dim, m2.img_encoder) == (4, 512, 16, 16, "vgg16")


def test_explainer_classes_exist_with_reference_signatures():
    import inspect
    from lrp_imagecaptioning_b200 import explainers as E
    for name in ("ExplainImgCaptioningAdaptiveAttention", "ExplainImgCaptioningAdaptiveAttentionGradient",
                 "ExplainImgCaptioningAdaptiveAttentionInputTimesGradient", "ExplainImgCaptioningAdaptiveAttentionGuidedGradcam",
                 "ExplainImgCaptioningGridTDModel", "ExplainImgCaptioningGridTDGradient",
                 "ExplainImgCaptioningGridTDGradientTimesInput", "ExplainImgCaptioningGridTDGuidedGradcam"):
        cls = getattr(E, name)
        assert list(inspect.signature(cls.__init__).parameters)[:5] == ["self", "model", "weight_path", "dataset_provider", "max_caption_length"]
        for meth in ("_forward_beam_search", "_explain_sentence", "_explain_CNN", "_beam_search"):
            assert hasattr(cls, meth)
    assert E.EPS == 0.01


def test_evaluation_argument_checks_need_no_gpu():
    """Mode / pooling names are validated on the host before anything touches the device."""
    from lrp_imagecaptioning_b200 import evaluation as EV
    m = np.zeros((1, 32, 32, 3), dtype=np.float32)
    with pytest.raises(ValueError):
        EV.heatmaps(m, "median")
    with pytest.raises(ValueError):
        EV.heatmaps(m, "mean", pooling="min")
    assert EV.THRESHOLDS == (0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)    # evaluate_bbox.py:251


def test_streamed_engine_split_is_a_partition():
    from lrp_imagecaptioning_b200.engine import StreamedEngine
    se = object.__new__(StreamedEngine)
    se.lanes = [None] * 3
    for n in (0, 1, 2, 3, 7, 64):
        b = se._split(n)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(2))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_keras_hdf5_reader_renames_datasets(monkeypatch, tmp_path):
    """keras_io walks the HDF5 datasets `<...>/<layer>/<tensor>:0` into the container's `<layer>/<tensor>` keys; h5py is
    absent here, so a stand-in module with the same File / visititems surface serves a synthetic model's tensors."""
    import sys
    import types
    from lrp_imagecaptioning_b200 import synth, keras_io
    from lrp_imagecaptioning_b200.model import CaptioningModel, _names
    from lrp_imagecaptioning_b200.encoder import ImageModel
    vgg = synth.vgg16_weights(1)
    dec = synth.decoder_weights("gridtd", V=50, H=64, E=64, D=512, seed=2)
    store = {}
    for (k, b), name in zip(vgg, ImageModel.layer_names):                      # VGG16 nested as a sub-model
        store["model_weights/vgg16/%s/kernel:0" % name] = k
        store["model_weights/vgg16/%s/bias:0" % name] = b
    # dataset names as Keras writes them for the reference's layers -- hard-coded from the reference's add_weight calls
    # (models/model.py:702-724: name='{}_W_va'.format(self.name) ...; the wrapped LSTM's tensors sit under the wrapper)
    GL = "external_bottom_up_attention_adaptive_1"
    real = {
        "image_features_w": "image_features/image_features/kernel:0", "image_features_b": "image_features/image_features/bias:0",
        "global_w": "global_img_feature/global_img_feature/kernel:0", "global_b": "global_img_feature/global_img_feature/bias:0",
        "embedding": "embedding_1/embedding_1/embeddings:0", "output_w": "output/output/kernel:0", "output_b": "output/output/bias:0",
        "lang_wi": GL + "/" + GL + "/kernel:0", "lang_wh": GL + "/" + GL + "/recurrent_kernel:0", "lang_b": GL + "/lstm_1/bias:0",
        "td_wi": GL + "/" + GL + "/" + GL + "_top_down_lstm_weight_i:0", "td_wh": GL + "/" + GL + "/" + GL + "_top_down_lstm_weight_h:0",
        "td_b": GL + "/" + GL + "/" + GL + "_top_down_lstm_weight_bias:0",
        "W_va": GL + "/" + GL + "/" + GL + "_W_va:0", "W_ha": GL + "/" + GL + "/" + GL + "_W_ha:0", "W_a": GL + "/" + GL + "/" + GL + "_W_a:0",
        "W_x": GL + "/" + GL + "/" + GL + "_W_x:0", "W_h": GL + "/" + GL + "/" + GL + "_W_h:0", "W_s": GL + "/" + GL + "/" + GL + "_W_s:0",
    }
    assert sorted(real) == sorted(_names("gridtd"))
    for ours, ds in real.items():
        store["model_weights/" + ds] = np.asarray(dec[ours])

    class FakeFile(object):
        def __init__(self, path, mode="r"):
            assert mode == "r"
        def __enter__(self):
            return self
        def __exit__(self, *a):
            return False
        def visititems(self, fn):
            fn("model_weights", object())                                       # a group: no shape / dtype
            for name, arr in store.items():
                fn(name, arr)
    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=FakeFile))
    w = keras_io.read_keras_hdf5("keras_model.hdf5")
    assert "block3_conv2/kernel" in w and "embedding_1/embeddings" in w and all(v.dtype == np.float32 for v in w.values())
    assert keras_io._key("model_weights/output/output/bias:0") == "output/bias"
    AL = "external_attention_rnn_wrapper_local_attention_v3_1"
    assert keras_io._key("model_weights/%s/%s/%s_Wv:0" % (AL, AL, AL)) == AL + "/Wv"
    assert keras_io._key("%s/lstm_2/recurrent_kernel:0" % AL) == AL + "/recurrent_kernel"
    # round trip through the container: hdf5 path -> same tensors as the synthetic source
    m = CaptioningModel("gridtd", synth.vgg16_weights(9), synth.decoder_weights("gridtd", V=50, H=64, E=64, D=512, seed=8),
                        image_hw=32)
    m.load_weights("keras_model.hdf5")
    assert np.array_equal(m.vgg[4][0], vgg[4][0]) and np.array_equal(m.dec["td_wh"], dec["td_wh"])
    names = keras_io.convert("keras_model.hdf5", str(tmp_path / "w.npz"))
    assert len(names) == len(store) and np.array_equal(np.load(str(tmp_path / "w.npz"))["output/kernel"], dec["output_w"])
    monkeypatch.delitem(sys.modules, "h5py")
    monkeypatch.setitem(sys.modules, "h5py", None)                              # import h5py -> ImportError
    with pytest.raises(ImportError):
        keras_io.read_keras_hdf5("keras_model.hdf5")


def test_conv_tile_chooser():
    """The accumulator tile of the generic tcgen05 kernel (csrc/tc_conv.cu: tc_conv_tile; host-only ABI call): at most 128
    rows, divides the map, and on the 14 / 28 / 56-wide VGG maps spans 9 or 16 words (126 / 128 rows instead of 98 / 112)."""
    from lrp_imagecaptioning_b200 import _lib
    for n in (1, 2, 5, 9, 20, 64, 320):
        for hw in (1, 2, 4, 7, 8, 14, 16, 28, 56, 112, 224):
            tw, th, ti = _lib.debug_conv_tile(n, hw, hw)
            assert 1 <= tw * th * ti <= 128 and 1 <= ti <= n
            if hw % 16 == 0 and hw >= 16:
                assert (tw, th, ti) == (16, 8, 1)                  # 16 x 8 pixels of one item
            else:
                assert hw % tw == 0 and hw % th == 0               # whole tiles only (no partially filled TMA boxes)
    assert _lib.debug_conv_tile(320, 14, 14) == (14, 1, 9)
    assert _lib.debug_conv_tile(320, 28, 28) == (14, 1, 9)
    assert _lib.debug_conv_tile(320, 56, 56) == (8, 1, 16)          # 56 = 7 x 8: all 128 rows
    assert _lib.debug_conv_tile(1, 14, 14) == (14, 7, 1)            # a single item: the old one-item tile
    assert _lib.debug_conv_tile(1, 28, 28) == (28, 4, 1)
    # fewest tiles wins: never more tiles than the one-item choice
    for n in (3, 20, 64, 320):
        for hw in (14, 28, 56):
            tw, th, ti = _lib.debug_conv_tile(n, hw, hw)
            tw1, th1, _ = _lib.debug_conv_tile(1, hw, hw)
            tiles = -(-n // ti) * (hw // tw) * (hw // th)
            assert tiles <= n * (hw // tw1) * (hw // th1)
