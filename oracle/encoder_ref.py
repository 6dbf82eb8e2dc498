"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the encoder stage. Never imported by the product.

PARITY UNPINNED at the TensorFlow boundary: the reference runs this stage as a Keras/TF-1.x graph
built by its modified iNNvestigate, and neither TensorFlow nor Keras can be installed here, and the
reference ships no golden vectors for it.  This file therefore restates, in torch-CPU float32, what
that graph computes for the VGG16 conv/max-pool chain (input_1 -> block5_conv3), following

  EpsilonRule          innvestigate/analyzer/relevance_based/relevance_rule.py:113-144
  ZRule                relevance_rule.py:74-98
  AlphaBetaRule (+ Alpha1Beta0 / Alpha2Beta1 / *IgnoreBias / ZPlus)   relevance_rule.py:216-368, 445-455
  ZPlusFastRule        relevance_rule.py:459-503
  SafeDivide / Divide  innvestigate/layers.py:436-461
  GradientWRT          innvestigate/layers.py:138-157 -> utils/keras/backend.py:45-60  (tf.gradients(Ys, Xs, grad_ys))
  max-pool / default   relevance_analyzer.py:459-480 (gradient routing)
  presets              relevance_analyzer.py:531-552, 578-623, 639-666, 678-692, 695-721
  'replace' seeding    analyzer/base.py:366-410, utils/keras/graph.py:898-900, 938
  Gradient / InputTimesGradient / GuidedBackprop   analyzer/gradient_based.py:101-172, 228-265

``torch.autograd.grad(z, x, s)`` is exactly ``tf.gradients(z, x, grad_ys=s)``.  What is checked
instead of golden vectors: relevance conservation for bias-free nets, the analyzer identities
iNNvestigate's own test helpers use (Z+ == alpha1beta0-ignore-bias for x>=0; gradient*input == LRP-Z
on bias-free ReLU nets), and shapes.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

POOL_AFTER = (1, 3, 6, 9)   # 0-based conv indices followed by a 2x2/2 max-pool (VGG16 to block5_conv3)
SAFE_FACTOR = 1e-7          # keras.backend.epsilon(), SafeDivide default factor


def _to_nchw(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).permute(0, 3, 1, 2).contiguous()


def _to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


def _conv(x, k_hwio, b):
    w = torch.from_numpy(np.ascontiguousarray(k_hwio)).permute(3, 2, 0, 1).contiguous()
    return Fn.conv2d(x, w, None if b is None else torch.from_numpy(np.ascontiguousarray(b)), padding=1)


def forward(images_nhwc, weights, n_layers=None):
    """Returns list of layer inputs xs[l] (NCHW tensors) and the final post-ReLU features (NHWC)."""
    x = _to_nchw(images_nhwc)
    xs = []
    n_layers = len(weights) if n_layers is None else n_layers
    for l in range(n_layers):
        k, b = weights[l]
        xs.append(x)
        x = torch.relu(_conv(x, k, b))
        if l in POOL_AFTER:
            xs.append(x)
            x = Fn.max_pool2d(x, 2, 2)
    return xs, _to_nhwc(x)


def features(images_nhwc, weights):
    return forward(images_nhwc, weights)[1]


def _safe_div(a, b):
    return a / (b + (b == 0).to(b.dtype) * SAFE_FACTOR)


def _grad(fn, x, s):
    x = x.detach().requires_grad_(True)
    z = fn(x)
    return torch.autograd.grad(z, x, s)[0]


def _rule_eps(x, k, b, R, eps, bias=True):
    f = lambda v: _conv(v, k, b if bias else None)
    z = f(x)
    s = R / (z + ((z >= 0).to(z.dtype) * 2 - 1) * eps)
    return x * _grad(f, x, s)


def _rule_z(x, k, b, R, bias=True):
    f = lambda v: _conv(v, k, b if bias else None)
    return x * _grad(f, x, _safe_div(R, f(x)))


def _rule_alphabeta(x, k, b, R, alpha, beta, bias=True):
    kp, kn = k * (k >= 0), k * (k < 0)
    bp = (b * (b >= 0)) if bias else None
    bn = (b * (b < 0)) if bias else None
    xp, xn = x * (x >= 0).to(x.dtype), x * (x < 0).to(x.dtype)

    def f(k1, b1, k2, b2):
        f1 = lambda v: _conv(v, k1, b1)
        f2 = lambda v: _conv(v, k2, b2)
        s = _safe_div(R, f1(xp) + f2(xn))
        return xp * _grad(f1, xp, s) + xn * _grad(f2, xn, s)

    act = f(kp, bp, kn, bn)
    if beta:
        inh = f(kn, bn, kp, bp)
        return alpha * act - beta * inh
    return act


def _rule_zplus_fast(x, k, b, R):
    kp = k * (k > 0)
    f = lambda v: _conv(v, kp, None)
    return x * _grad(f, x, _safe_div(R, f(x)))


def _infer_alpha_beta(alpha, beta):
    """relevance_based/utils.py:72-129."""
    if alpha is None and beta is None:
        raise ValueError("Neither alpha or beta were given")
    if alpha is None:
        alpha = beta + 1
    if beta is None:
        beta = alpha - 1
    if alpha < 1 or beta < 0 or abs((alpha - beta) - 1) > 1e-12:
        raise ValueError("alpha >= 1, beta >= 0 and alpha - beta = 1 are required")
    return alpha, beta


def analyze(method, images_nhwc, R_head_nhwc, weights, epsilon=1e-7, alpha=None, beta=None, bias=True):
    """Relevance / gradient of the supplied head tensor at block5_conv3's output, at the input image.

    method: 'lrp.epsilon' | 'lrp.z' | 'lrp.alpha_beta' | 'lrp.alpha_1_beta_0' | 'lrp.alpha_2_beta_1' |
            'lrp.z_plus' | 'lrp.z_plus_fast' | 'lrp.sequential_preset_a' |
            'gradient' | 'input_t_gradient' | 'guided_backprop'
    """
    xs, _ = forward(images_nhwc, weights)
    R = _to_nchw(R_head_nhwc)
    img = xs[0]
    if method in ("gradient", "input_t_gradient", "guided_backprop"):
        j = len(xs) - 1
        for l in range(len(weights) - 1, -1, -1):
            k, b = weights[l]
            if l in POOL_AFTER:
                R = _grad(lambda v: Fn.max_pool2d(v, 2, 2), xs[j], R)
                j -= 1
            if method == "guided_backprop":
                R = torch.relu(R)
            R = _grad(lambda v: torch.relu(_conv(v, k, b)), xs[j], R)
            j -= 1
        if method == "input_t_gradient":
            R = R * img
        return _to_nhwc(R)

    if method == "lrp.sequential_preset_a":      # conv layers: Alpha1Beta0Rule (bias on); no Dense in the encoder
        method, alpha, beta, bias = "lrp.alpha_beta", 1, 0, True
    elif method == "lrp.alpha_1_beta_0":
        method, alpha, beta, bias = "lrp.alpha_beta", 1, 0, True
    elif method == "lrp.alpha_2_beta_1":
        method, alpha, beta, bias = "lrp.alpha_beta", 2, 1, True
    elif method == "lrp.z_plus":
        method, alpha, beta, bias = "lrp.alpha_beta", 1, 0, False
    if method == "lrp.alpha_beta":
        alpha, beta = _infer_alpha_beta(alpha, beta)
    if method == "lrp.epsilon" and not epsilon > 0:
        raise ValueError("epsilon must be > 0")
    j = len(xs) - 1
    for l in range(len(weights) - 1, -1, -1):
        k, b = weights[l]
        if l in POOL_AFTER:
            R = _grad(lambda v: Fn.max_pool2d(v, 2, 2), xs[j], R)
            j -= 1
        x = xs[j]
        j -= 1
        if method == "lrp.epsilon":
            R = _rule_eps(x, k, b, R, epsilon, bias)
        elif method == "lrp.z":
            R = _rule_z(x, k, b, R, bias)
        elif method == "lrp.alpha_beta":
            R = _rule_alphabeta(x, k, b, R, alpha, beta, bias)
        elif method == "lrp.z_plus_fast":
            R = _rule_zplus_fast(x, k, b, R)
        else:
            raise ValueError("unknown method %r" % (method,))
    return _to_nhwc(R)
