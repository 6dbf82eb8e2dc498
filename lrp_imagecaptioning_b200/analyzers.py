"""Drop-in for the iNNvestigate analyzer entry points the captioning explainers use.

Mirrors (names, kwargs, error behaviour) of the reference's modified iNNvestigate for a VGG16 `_image_model`:
  create_analyzer / analyzers registry   innvestigate/analyzer/__init__.py:35-99
  LRP presets                            innvestigate/analyzer/relevance_based/relevance_analyzer.py:531-721
  parameter checks                       innvestigate/analyzer/relevance_based/utils.py:52-129
  Gradient / InputTimesGradient / GuidedBackprop   innvestigate/analyzer/gradient_based.py:101-172, 228-265
  'replace' neuron-selection mode        innvestigate/analyzer/base.py:327-334, 366-410, 478-520

`analyze([X, R])`: X = images [N, hw, hw, 3]; R = head tensor [N, hw/16, hw/16, 512] that *replaces* the
head relevance at block5_conv3's output; returns a numpy array shaped like X (relevance in the model's BGR
channel order, as the reference does -- callers flip with postprocess(..., 'BGRtoRGB')).
`analyze_batch(X, img_index, R)` is the batched form (many words per image; per-image work done once).
"""
import numpy as np

from . import _lib
from .encoder import ImageModel, RuleSpec


class NotAnalyzeableModelException(Exception):
    """innvestigate/analyzer/base.py:33-35."""


def _check_epsilon(epsilon, caller):
    if epsilon <= 0:
        raise ValueError("Constructor call to {} : Parameter epsilon must be > 0 but was {}".format(
            caller.__class__.__name__, epsilon))
    return epsilon


def _infer_alpha_beta(alpha, beta, caller):
    head = "Constructor call to {} : ".format(caller.__class__.__name__)
    if alpha is None and beta is None:
        raise ValueError(head + "Neither alpha or beta were given")
    if alpha is not None and alpha < 1:
        raise ValueError(head + "Passed parameter alpha invalid. Expecting alpha >= 1 but was {}".format(alpha))
    if beta is not None and beta < 0:
        raise ValueError(head + "Passed parameter beta invalid. Expecting beta >= 0 but was {}".format(beta))
    if alpha is None:
        alpha = beta + 1
    if beta is None:
        beta = alpha - 1
    if alpha - beta != 1:
        raise ValueError(head + "Condition alpha - beta = 1 not fulfilled. alpha={} ; beta={} -> alpha - beta = {}".format(
            alpha, beta, alpha - beta))
    return alpha, beta


class AnalyzerBase(object):
    """Common part: 'replace' mode only (the other modes are broken in the reference fork)."""

    def __init__(self, model, neuron_selection_mode="replace", allow_lambda_layers=False, **kwargs):
        if neuron_selection_mode not in ["max_activation", "index", "all", "replace"]:
            raise ValueError("neuron_selection parameter is not valid.")
        if neuron_selection_mode != "replace":
            raise NotImplementedError("only neuron_selection_mode='replace' is supported (the reference fork seeds "
                                      "the backward pass with model.inputs[1] unconditionally, graph.py:898-900)")
        if not isinstance(model, ImageModel):
            raise NotAnalyzeableModelException("model must be an lrp_imagecaptioning_b200.encoder.ImageModel "
                                               "(VGG16 input_1 -> block5_conv3)")
        if kwargs:
            raise TypeError("unexpected keyword arguments: %s" % sorted(kwargs))
        self._model = model
        self._neuron_selection_mode = neuron_selection_mode

    def _rule(self):
        raise NotImplementedError

    def analyze(self, X):
        if not isinstance(X, (list, tuple)) or len(X) != 2:
            raise ValueError("'replace' mode expects X = [images, head_relevance]")
        imgs, R = X
        imgs = np.asarray(imgs, dtype=np.float32)
        n = imgs.shape[0]
        return self.analyze_batch(imgs, np.arange(n, dtype=np.int32), R).cpu().numpy()

    def analyze_batch(self, images, img_index, R_head):
        """images [N,...]; word w uses image img_index[w] and head tensor R_head[w]. Returns a CUDA tensor."""
        rule = self._rule()
        self._model.forward(images, rule)
        return self._model.relevance(img_index, R_head)

    def analyze_resident(self, img_index, R_head):
        """Backward only: reuses the state of the last forward (must have used this analyzer's rule)."""
        st = self._model._state
        if st is None or st[0] != self._rule().key():
            raise _lib.LrpcapError(-5, "encoder state was built for a different rule; call analyze_batch")
        return self._model.relevance(img_index, R_head)


class LRPEpsilon(AnalyzerBase):
    def __init__(self, model, epsilon=1e-7, bias=True, *args, **kwargs):
        self._epsilon = _check_epsilon(epsilon, self)
        self._bias = bias
        super(LRPEpsilon, self).__init__(model, *args, **kwargs)

    def _rule(self):
        return RuleSpec(_lib.RULE_EPSILON, epsilon=self._epsilon, bias=self._bias)


class LRPEpsilonIgnoreBias(LRPEpsilon):
    def __init__(self, model, epsilon=1e-7, *args, **kwargs):
        super(LRPEpsilonIgnoreBias, self).__init__(model, epsilon=epsilon, bias=False, *args, **kwargs)


class LRPZ(AnalyzerBase):
    def __init__(self, model, bias=True, *args, **kwargs):
        self._bias = bias
        super(LRPZ, self).__init__(model, *args, **kwargs)

    def _rule(self):
        return RuleSpec(_lib.RULE_Z, bias=self._bias)


class LRPZIgnoreBias(LRPZ):
    def __init__(self, model, *args, **kwargs):
        super(LRPZIgnoreBias, self).__init__(model, bias=False, *args, **kwargs)


class LRPAlphaBeta(AnalyzerBase):
    def __init__(self, model, alpha=None, beta=None, bias=True, *args, **kwargs):
        self._alpha, self._beta = _infer_alpha_beta(alpha, beta, self)
        self._bias = bias
        super(LRPAlphaBeta, self).__init__(model, *args, **kwargs)

    def _rule(self):
        return RuleSpec(_lib.RULE_ALPHA_BETA, alpha=self._alpha, beta=self._beta, bias=self._bias)


class LRPAlpha2Beta1(LRPAlphaBeta):
    def __init__(self, model, *args, **kwargs):
        super(LRPAlpha2Beta1, self).__init__(model, alpha=2, beta=1, bias=True, *args, **kwargs)


class LRPAlpha2Beta1IgnoreBias(LRPAlphaBeta):
    def __init__(self, model, *args, **kwargs):
        super(LRPAlpha2Beta1IgnoreBias, self).__init__(model, alpha=2, beta=1, bias=False, *args, **kwargs)


class LRPAlpha1Beta0(LRPAlphaBeta):
    def __init__(self, model, *args, **kwargs):
        super(LRPAlpha1Beta0, self).__init__(model, alpha=1, beta=0, bias=True, *args, **kwargs)


class LRPAlpha1Beta0IgnoreBias(LRPAlphaBeta):
    def __init__(self, model, *args, **kwargs):
        super(LRPAlpha1Beta0IgnoreBias, self).__init__(model, alpha=1, beta=0, bias=False, *args, **kwargs)


class LRPZPlus(LRPAlpha1Beta0IgnoreBias):
    pass


class LRPZPlusFast(AnalyzerBase):
    def _rule(self):
        return RuleSpec(_lib.RULE_ZPLUS_FAST)


class LRPSequentialPresetA(AnalyzerBase):
    """relevance_analyzer.py:695-721: Dense -> EpsilonRule(epsilon), Conv -> Alpha1Beta0Rule.  The VGG16 encoder
    has no Dense layer, so epsilon is validated and otherwise unused."""

    def __init__(self, model, epsilon=1e-1, *args, **kwargs):
        self._epsilon = _check_epsilon(epsilon, self)
        super(LRPSequentialPresetA, self).__init__(model, *args, **kwargs)

    def _rule(self):
        return RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True)


class Gradient(AnalyzerBase):
    def __init__(self, model, postprocess=None, **kwargs):
        if postprocess not in [None, "abs", "square"]:
            raise ValueError("Parameter 'postprocess' must be either None, 'abs', or 'square'.")
        self._postprocess = postprocess
        super(Gradient, self).__init__(model, **kwargs)

    def _rule(self):
        return RuleSpec(_lib.RULE_GRADIENT)

    def analyze_batch(self, images, img_index, R_head):
        out = super(Gradient, self).analyze_batch(images, img_index, R_head)
        if self._postprocess == "abs":
            out = out.abs()
        elif self._postprocess == "square":
            out = out * out
        return out


class InputTimesGradient(Gradient):
    def __init__(self, model, **kwargs):
        super(InputTimesGradient, self).__init__(model, **kwargs)

    def _rule(self):
        return RuleSpec(_lib.RULE_INPUT_T_GRADIENT)


class GuidedBackprop(AnalyzerBase):
    def _rule(self):
        return RuleSpec(_lib.RULE_GUIDED_BACKPROP)


analyzers = {
    "gradient": Gradient,
    "input_t_gradient": InputTimesGradient,
    "guided_backprop": GuidedBackprop,
    "lrp.z": LRPZ,
    "lrp.z_IB": LRPZIgnoreBias,
    "lrp.epsilon": LRPEpsilon,
    "lrp.epsilon_IB": LRPEpsilonIgnoreBias,
    "lrp.alpha_beta": LRPAlphaBeta,
    "lrp.alpha_2_beta_1": LRPAlpha2Beta1,
    "lrp.alpha_2_beta_1_IB": LRPAlpha2Beta1IgnoreBias,
    "lrp.alpha_1_beta_0": LRPAlpha1Beta0,
    "lrp.alpha_1_beta_0_IB": LRPAlpha1Beta0IgnoreBias,
    "lrp.z_plus": LRPZPlus,
    "lrp.z_plus_fast": LRPZPlusFast,
    "lrp.sequential_preset_a": LRPSequentialPresetA,
}


def create_analyzer(name, model, **kwargs):
    """innvestigate/analyzer/__init__.py:88-99."""
    return analyzers[name](model, **kwargs)
