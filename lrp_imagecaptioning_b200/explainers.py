"""Drop-in for the explainer classes of the reference's models/explainers.py, backed by the CUDA library.

Same class names, constructor signature `(model, weight_path, dataset_provider, max_caption_length)`, method
names and return shapes as /root/reference/models/explainers.py (:22-189, :260-278, :667-671, :880-949, :995-1019,
:1322-1325, :1585-1653); the two-phase protocol (`_forward_beam_search` then `_explain_*`) and the attributes callers
read (`.caption`, `.attention`, `.beta`, `.r_words`) are kept.  `model` is a model.CaptioningModel instead of a Keras
model wrapper.  Callers that pass `rule='eps'` (exaimin_word.py:93,189) are accepted; the decoder only has the
epsilon rule.

Differences a maintainer should know (all switchable or documented in DESIGN.md):
  * `_explain_sentence` runs every word of the sentence in one batched launch sequence instead of a Python loop.
  * `_beam_search` ranks hypotheses with the decoder's last-step logits (one batched teacher-forced forward per
    round for all images x beams) instead of re-running the whole Keras model incl. VGG16 per round.
"""
import numpy as np

from . import _lib
from .analyzers import LRPSequentialPresetA, Gradient, InputTimesGradient, GuidedBackprop
from .decoder import DecoderEngine

EPS = 0.01
ALPHA = 1
BETA = 0


class _DefaultPreprocessor(object):
    SOS_TOKEN_LABEL_ENCODED = 1
    EOS_TOKEN_LABEL_ENCODED = 2


class ExplainImgCaptioningAttentionModel(object):
    _decoder_kind = None

    def __init__(self, model, weight_path, dataset_provider, max_caption_length):
        if self._decoder_kind is not None and model.kind != self._decoder_kind:
            raise ValueError("%s needs a %r model, got %r" % (type(self).__name__, self._decoder_kind, model.kind))
        self._model = model
        if weight_path:
            model.load_weights(weight_path)
        self._image_model = model.image_model
        self._img_encoder = model.img_encoder
        self._CNN_explainer = LRPSequentialPresetA(self._image_model, epsilon=EPS, neuron_selection_mode="replace")
        # the reference always gets a DatasetProvider; without one (dataset_provider=None) the tokenizer's conventional
        # ids stand in: SOS = 1, EOS = 2 (models/preprocessors.py: the first two fitted tokens)
        self._preprocessor = dataset_provider.caption_preprocessor if dataset_provider is not None else _DefaultPreprocessor()
        self._dataset_provider = dataset_provider
        self._max_caption_length = max_caption_length
        self._hidden_dim = model._hidden_dim
        self._embedding_dim = model._embedding_dim
        self.L = model.L
        self.D = model.D
        self._weight_path = (weight_path or "").strip("hdf5")
        self._decoder = DecoderEngine(model.dec, sos=self._preprocessor.SOS_TOKEN_LABEL_ENCODED, device=model.device)
        self._img = None

    # ---- caption generation (explainers.py:51-120; inference.py:267-315 BatchNLargest / NLargest / Caption)
    def _beam_search(self, X, beam_size):
        """Beam search over the captioner's logits (the Keras model's: (h + c_hat) W_o + b for both kinds), with the
        reference's bookkeeping: hypotheses start as [SOS, EOS]; each round extends every hypothesis by its `beam_size`
        best words, keeps the best `beam_size` per image; a hypothesis extended by EOS records its parent as complete.
        Returns, per beam rank, the token list without SOS (`[words..., EOS]`) -- the structure `explain_image.py:41-43`
        indexes as `_beam_search(X, 3)[0]` for a single image."""
        import heapq
        _, imgs = X
        imgs = np.asarray(imgs, dtype=np.float32)
        B = imgs.shape[0]
        pre = self._preprocessor
        SOS, EOS = pre.SOS_TOKEN_LABEL_ENCODED, pre.EOS_TOKEN_LABEL_ENCODED
        self._image_model.forward(imgs, self._CNN_explainer._rule())
        feats = self._image_model.features().reshape(B, self.L, self.D)
        if getattr(self, "_decoder_keras", None) is None:
            self._decoder_keras = DecoderEngine(self._model.dec, sos=SOS, keras_logits=True, device=self._model.device)

        def push(heap, item):
            if len(heap) < beam_size:
                heapq.heappush(heap, item)
            else:
                heapq.heappushpop(heap, item)
        partial = [[(0.0, [SOS, EOS])] for _ in range(B)]
        complete = [[] for _ in range(B)]
        for _ in range(self._max_caption_length):
            hyps = [(b, lp, sent) for b in range(B) for (lp, sent) in sorted(partial[b], reverse=True)]
            tokens = np.asarray([sent[1:-1] + [SOS] for (_, _, sent) in hyps], dtype=np.int32)   # last token is never read
            self._decoder_keras.forward(feats[[b for (b, _, _) in hyps]], tokens)
            logits = self._decoder_keras.last_logits()
            logp = logits - np.max(logits, axis=-1, keepdims=True)
            logp = logp - np.log(np.sum(np.exp(logp), axis=-1, keepdims=True))
            partial = [[] for _ in range(B)]
            for (b, lp_prev, sent), row in zip(hyps, logp):
                top = np.argpartition(row, -beam_size)[-beam_size:]
                for w in top:
                    word = int(w) + 1                       # model index -> tokenizer id
                    lp = float(row[w] + lp_prev)
                    push(partial[b], (lp, sent[:-1] + [word, sent[-1]]))
                    if word == EOS:
                        push(complete[b], (lp, sent))
        results = []
        for rank in range(beam_size):
            row = []
            for b in range(B):
                top_p = sorted(partial[b], reverse=True)
                top_c = sorted(complete[b], reverse=True)
                cap = top_c[rank] if rank < len(top_c) else (top_p[rank] if rank < len(top_p) else None)
                row.append(cap[1][1:] if cap is not None else None)
            results.append(row[0] if B == 1 else row)
        return results

    # ---- forward with stored state (explainers.py:370-436, 1092-1178)
    def _forward_beam_search(self, X, beam_search_captions):
        _, img_input = X
        img = np.ascontiguousarray(np.asarray(img_input, dtype=np.float32))
        if img.ndim != 4 or img.shape[0] != 1:
            raise ValueError("expected one image [1, hw, hw, 3]")
        self.caption = [int(c) for c in beam_search_captions]
        self._img = img
        self._image_model.forward(img, self._CNN_explainer._rule())
        self._img_feature_input = None
        self._decoder.forward(self._image_model.features(), np.asarray([self.caption], dtype=np.int32))
        al, be = self._decoder.attention()
        self.attention = al[0]
        self.beta = be[0][:, None]

    def _features_numpy(self):
        if self._img_feature_input is None:
            self._img_feature_input = self._image_model.features().cpu().numpy()[0].reshape(self.L, self.D)
        return self._img_feature_input

    # ---- decoder relevance (explainers.py:537-666, 1180-1321)
    def _explain_lstm_single_word_sequence(self, t=0, rule=None):
        if t > len(self.caption) or t < 1:
            raise NotImplementedError("index out of range of captions")
        R, rw, att = self._decoder.relevance([0], [t])
        self._set_r_words(rw[0], t)
        side = int(np.sqrt(self.L))
        return R.cpu().numpy().reshape(1, side, side, self.D), att[0]

    def _set_r_words(self, row, t):
        self.r_words = np.array(row[:t])

    def _explain_sentence(self, rule=None):
        n = len(self.caption) - 1            # the last token is skipped (explainers.py:186)
        if n <= 0:
            return [], self.attention[1:-1]
        R, rw, _ = self._decoder.relevance([0] * n, list(range(1, n + 1)), want_attention=False)
        R = R.cpu().numpy()
        side = int(np.sqrt(self.L))
        self._set_r_words(rw[n - 1], n)
        return [R[i].reshape(1, side, side, self.D) for i in range(n)], self.attention[1:-1]

    # ---- encoder relevance (explainers.py:179-181)
    def _explain_CNN(self, X, relevance_value):
        X = np.asarray(X, dtype=np.float32)
        R = np.asarray(relevance_value, dtype=np.float32)
        st = self._image_model._state
        if (self._img is not None and st is not None and st[0] == self._CNN_explainer._rule().key()
                and st[1] == 1 and X.shape == self._img.shape and np.array_equal(X, self._img)):
            return self._CNN_explainer.analyze_resident(np.zeros(R.shape[0], dtype=np.int32), R).cpu().numpy()
        return self._CNN_explainer.analyze([X, R])

    def explain_sentence_to_pixels(self):
        """Batched convenience: decoder + encoder relevance for words 1..len-1 -> [n, hw, hw, 3] numpy."""
        n = len(self.caption) - 1
        R, _, _ = self._decoder.relevance([0] * n, list(range(1, n + 1)), want_words=False, want_attention=False)
        fh = int(np.sqrt(self.L))
        return self._CNN_explainer.analyze_resident(np.zeros(n, dtype=np.int32), R.view(n, fh, fh, self.D)).cpu().numpy()


class _GradientMixin(object):
    """Manual BPTT with frozen attention (explainers.py:780-832, 1452-1532)."""

    def _lstm_decoder_backward(self, t):
        if t > len(self.caption) or t < 1:
            raise NotImplementedError("index out of range of captions")
        R, rw = self._decoder.backward([0], [t])
        self.r_words = np.array(rw[0][:t])
        side = int(np.sqrt(self.L))
        return R.cpu().numpy().reshape(1, side, side, self.D)

    def _explain_sentence(self, rule=None):
        n = len(self.caption) - 1
        if n <= 0:
            return []
        R, rw = self._decoder.backward([0] * n, list(range(1, n + 1)))
        R = R.cpu().numpy()
        self.r_words = np.array(rw[n - 1][:n])
        side = int(np.sqrt(self.L))
        return [R[i].reshape(1, side, side, self.D) for i in range(n)]


class ExplainImgCaptioningAdaptiveAttention(ExplainImgCaptioningAttentionModel):
    _decoder_kind = "adaptive"

    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningAdaptiveAttention, self).__init__(model, weight_path, dataset_provider, max_caption_length)

    def _set_r_words(self, row, t):
        self.r_words = np.array(row[:max(t - 1, 0)])   # normalised, first entry dropped (explainers.py:660-665)


class ExplainImgCaptioningAdaptiveAttentionGradient(_GradientMixin, ExplainImgCaptioningAdaptiveAttention):
    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningAdaptiveAttentionGradient, self).__init__(model, weight_path, dataset_provider, max_caption_length)
        self._CNN_explainer = Gradient(self._image_model, neuron_selection_mode="replace")


class ExplainImgCaptioningAdaptiveAttentionInputTimesGradient(ExplainImgCaptioningAdaptiveAttentionGradient):
    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningAdaptiveAttentionInputTimesGradient, self).__init__(model, weight_path, dataset_provider, max_caption_length)
        self._CNN_explainer = InputTimesGradient(self._image_model, neuron_selection_mode="replace")


class _GuidedGradcamMixin(object):
    """explainers.py:930-949 / 1634-1653: guided backprop x Grad-CAM."""

    def _explain_CNN(self, X, relevance_value):
        from .gradcam import grad_cam
        gradcamp = grad_cam(self._features_numpy(), np.asarray(relevance_value)[0], self.L, self.D)
        guided = super(_GuidedGradcamMixin, self)._explain_CNN(X, relevance_value)
        return (guided[0] * gradcamp[..., np.newaxis])[np.newaxis, :]

    def grad_cam(self, img_feature, grads):
        from .gradcam import grad_cam
        return grad_cam(img_feature, grads, self.L, self.D)


class ExplainImgCaptioningAdaptiveAttentionGuidedGradcam(_GuidedGradcamMixin, ExplainImgCaptioningAdaptiveAttentionGradient):
    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningAdaptiveAttentionGuidedGradcam, self).__init__(model, weight_path, dataset_provider, max_caption_length)
        self._CNN_explainer = GuidedBackprop(self._image_model, neuron_selection_mode="replace")


class ExplainImgCaptioningGridTDModel(ExplainImgCaptioningAttentionModel):
    _decoder_kind = "gridtd"

    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningGridTDModel, self).__init__(model, weight_path, dataset_provider, max_caption_length)


class ExplainImgCaptioningGridTDGradient(_GradientMixin, ExplainImgCaptioningGridTDModel):
    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningGridTDGradient, self).__init__(model, weight_path, dataset_provider, max_caption_length)
        self._CNN_explainer = Gradient(self._image_model, neuron_selection_mode="replace")


class ExplainImgCaptioningGridTDGradientTimesInput(ExplainImgCaptioningGridTDGradient):
    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningGridTDGradientTimesInput, self).__init__(model, weight_path, dataset_provider, max_caption_length)
        self._CNN_explainer = InputTimesGradient(self._image_model, neuron_selection_mode="replace")


class ExplainImgCaptioningGridTDGuidedGradcam(_GuidedGradcamMixin, ExplainImgCaptioningGridTDGradient):
    def __init__(self, model, weight_path, dataset_provider, max_caption_length=20):
        super(ExplainImgCaptioningGridTDGuidedGradcam, self).__init__(model, weight_path, dataset_provider, max_caption_length)
        self._CNN_explainer = GuidedBackprop(self._image_model, neuron_selection_mode="replace")
