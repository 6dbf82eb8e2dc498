"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/decoder_<kind>.npz by executing the REFERENCE's own NumPy decoder
(/root/reference/models/explainers.py, imported under the Keras stub of oracle/refstub.py) on seeded synthetic weights.
Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/make_golden.py

Stored per kind: the weights/inputs and, for every word t, the reference's feature relevance, attention, r_words
(LRP classes) and decoder gradient (Gradient classes), plus the forward logits.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lrp_imagecaptioning_b200 import synth  # noqa: E402
from oracle import refstub  # noqa: E402

CFG = dict(V=30, H=16, E=16, D=24, L=9, T=5)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for kind in ("adaptive", "gridtd"):
        dec = synth.decoder_weights(kind, V=CFG["V"], H=CFG["H"], E=CFG["E"], D=CFG["D"], seed=41)
        F = synth.features(1, L=CFG["L"], D=CFG["D"], seed=42)[0]
        cap = [int(c) for c in synth.captions(1, CFG["T"], CFG["V"], seed=43)[0]]
        store = {"F": F, "caption": np.array(cap, dtype=np.int32)}
        for k, v in dec.items():
            if isinstance(v, np.ndarray):
                store["w_" + k] = v
        ref = refstub.make_reference_explainer(kind, "lrp", dec, F)
        ref._forward_beam_search((None, None), cap)
        store["logits"] = np.asarray(ref.caption_preds, dtype=np.float64)
        for t in range(1, CFG["T"] + 1):
            r, att = ref._explain_lstm_single_word_sequence(t)
            store["lrp_R_%d" % t] = np.asarray(r)
            store["lrp_att_%d" % t] = np.asarray(att, dtype=np.float64)
            store["lrp_rwords_%d" % t] = np.asarray(ref.r_words, dtype=np.float64)
        refg = refstub.make_reference_explainer(kind, "gradient", dec, F)
        refg._forward_beam_search((None, None), cap)
        for t in range(1, CFG["T"] + 1):
            store["grad_R_%d" % t] = np.asarray(refg._lstm_decoder_backward(t))
            store["grad_rwords_%d" % t] = np.asarray(refg.r_words, dtype=np.float64)
        path = os.path.join(out_dir, "decoder_%s.npz" % kind)
        np.savez_compressed(path, **store)
        print(path, os.path.getsize(path), "bytes")


def lrp_inference_case(kind):
    """Shared by the fixture generator and the tests (weights are re-synthesised from the seeds, not stored)."""
    vgg = synth.vgg16_weights(0)
    dec = synth.decoder_weights(kind, V=30, H=16, E=16, D=512, seed=1)
    imgs = synth.images(2, 32, 2)
    yp = np.random.default_rng(3).standard_normal((2, 4, 30))
    yp[..., -1] = -100.0     # the reference overflows its (V,) buffer when the arg-max is the last index (quirk B10)
    yp[1, 2, 1] = 50.0       # EOS (id 2) predicted at position 3 of sample 1
    word_of = {i: "w%d" % i for i in range(1, 31)}
    word_of[int(np.argmax(yp[0, 1]) + 1)] = "the"    # a stop word
    return vgg, dec, imgs, yp, word_of


def main_lrp_inference():
    out_dir = os.path.join(ROOT, "tests", "golden")
    _, M = refstub.load_reference()
    for kind in ("adaptive", "gridtd"):
        vgg, dec, imgs, yp, word_of = lrp_inference_case(kind)
        store = {"stop_words": np.array(sorted(set(M.STOP_WORDS)))}
        for mode in ("mean", "pos_mean", "quantile"):
            ref = refstub.make_reference_lrp_inference_layer(kind, dec, vgg, mode=mode, word_of=word_of)
            store[mode] = ref.call([np.zeros((2, 4), dtype=np.int32), imgs, yp.copy()])
        path = os.path.join(out_dir, "lrp_inference_%s.npz" % kind)
        np.savez_compressed(path, **store)
        print(path, os.path.getsize(path), "bytes")


def evaluation_case():
    """Seeded pixel maps, boxes and thresholds shared by the generator and the tests."""
    rng = np.random.default_rng(5)
    maps = (rng.standard_normal((3, 224, 224, 3)) * (rng.random((3, 224, 224, 1)) > 0.6)).astype(np.float32)
    boxes = [(0, 10, 20, 120, 200), (1, 100, 3, 224, 77), (2, 0, 0, 224, 224), (0, 50, 50, 51, 51)]
    thresholds = [0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9]
    return maps, boxes, thresholds


def reference_evaluation_objects():
    """The reference's evaluation classes (evaluate_bbox.py:39, exaimin_word.py:27) under the Keras stub, with the
    explainer replaced by a stand-in that returns a prepared pixel map; skimage's pyramid_expand (absent) is only used
    for the attention map, which this path does not check."""
    import importlib
    refstub.load_reference()
    eb = importlib.import_module("evaluate_bbox")
    ew = importlib.import_module("exaimin_word")
    eb.skimage.transform.pyramid_expand = lambda a, **k: np.zeros((224, 224))
    eb.K.image_data_format = lambda: "channels_last"

    class FakeExplainer(object):
        def __init__(self):
            self.map = None

        def _explain_lstm_single_word_sequence(self, t):
            return None, np.zeros(196)

        def _explain_CNN(self, img, rel):
            return self.map[None].copy()
    bb = object.__new__(eb.EvaluationBboxCOCO)
    bb._img_encoder, bb._color_conversion, bb._explainer = "vgg16", "BGRtoRGB", FakeExplainer()
    ex = object.__new__(ew.Explainer)
    return bb, ex


def main_evaluation():
    maps, boxes, thresholds = evaluation_case()
    bb, ex = reference_evaluation_objects()
    store = {"heat_negative": [], "ratios": [], "pool_max": [], "pool_ave": []}
    for i in range(maps.shape[0]):
        bb._explainer.map = maps[i]
        hm, _ = bb._get_explanation((None, None), 1)
        store["heat_negative"].append(hm)
        hp = np.mean(maps[i][..., ::-1], axis=-1)
        store["pool_max"].append(ex._max_pooling(hp))
        store["pool_ave"].append(ex._ave_pooling(hp))
    for (mi, x0, y0, x1, y1) in boxes:
        rel = store["heat_negative"][mi].copy()
        store["ratios"].append([bb._calculate_overlaped_pixels([x0, y0, x1, y1], rel, th) for th in thresholds])
    path = os.path.join(ROOT, "tests", "golden", "evaluation.npz")
    np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32 if k != "ratios" else np.float64) for k, v in store.items()})
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main_lrp_inference()
    main_evaluation()
