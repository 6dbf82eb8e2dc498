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
POOL_AFTER_VGG19 = (1, 3, 7, 11)   # VGG19 to block5_conv4 (16 convs)


def _pools(weights):
    return POOL_AFTER_VGG19 if len(weights) == 16 else POOL_AFTER
SAFE_FACTOR = 1e-7          # keras.backend.epsilon(), SafeDivide default factor


def _to_nchw(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).permute(0, 3, 1, 2).contiguous()


def _to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().numpy()


def _conv(x, k_hwio, b):
    w = torch.from_numpy(np.ascontiguousarray(k_hwio)).permute(3, 2, 0, 1).contiguous()
    return Fn.conv2d(x, w, None if b is None else torch.from_numpy(np.ascontiguousarray(b)), padding=1)


class Forced(object):
    """Discrete decisions taken from another implementation's forward pass (tests pin the oracle to them: every rule
    is discontinuous at max-pool arg-max ties, the gradient family and the Z rule also at ReLU kinks).

    routes: {conv layer index l in POOL_AFTER: uint8 [N, H/2, W/2, C]}, window position sy*2+sx the pool after layer l
            routes to (instead of torch's own arg-max);
    masks:  optional {conv layer index: bool [N, H, W, C]} = [z_l > 0] (instead of the oracle's own ReLU decision).
    """

    def __init__(self, routes=None, masks=None):
        self.routes = {} if routes is None else {int(l): torch.from_numpy(np.ascontiguousarray(r)).permute(0, 3, 1, 2).long()
                                                 for l, r in routes.items()}
        self.masks = {} if masks is None else {int(l): torch.from_numpy(np.ascontiguousarray(m).astype(np.float32)).permute(0, 3, 1, 2).contiguous()
                                               for l, m in masks.items()}

    def subset(self, idx):
        f = Forced()
        idx = torch.as_tensor(np.asarray(idx), dtype=torch.long)
        f.routes = {l: r[idx] for l, r in self.routes.items()}
        f.masks = {l: m[idx] for l, m in self.masks.items()}
        return f


def _pool(x, l, force=None):
    """2x2/2 max-pool after conv layer l; with forced routes the pooled value is x at the forced window position, so
    that autograd routes the cotangent there (what TF MaxPoolGrad does for the arg-max it found)."""
    if force is None or l not in force.routes:
        return Fn.max_pool2d(x, 2, 2)
    N, C, H, W = x.shape
    win = x.reshape(N, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(N, C, H // 2, W // 2, 4)
    return torch.gather(win, 4, force.routes[l].unsqueeze(-1)).squeeze(-1)


def _relu(z, l, force=None):
    if force is None or l not in force.masks:
        return torch.relu(z)
    return z * force.masks[l]


def forward(images_nhwc, weights, n_layers=None, force=None):
    """Returns list of layer inputs xs[l] (NCHW tensors) and the final post-ReLU features (NHWC)."""
    x = _to_nchw(images_nhwc)
    xs = []
    n_layers = len(weights) if n_layers is None else n_layers
    for l in range(n_layers):
        k, b = weights[l]
        xs.append(x)
        x = _relu(_conv(x, k, b), l, force)
        if l in _pools(weights):
            xs.append(x)
            x = _pool(x, l, force)
    return xs, _to_nhwc(x)


def pool_routes(images_nhwc, weights):
    """The oracle's own arg-max routes, same format as Forced.routes input: {l: uint8 [N, H/2, W/2, C]} (first maximum)."""
    xs, _ = forward(images_nhwc, weights)
    out, j = {}, 0
    for l in range(len(weights)):
        j += 1
        if l in _pools(weights):
            x = xs[j]
            j += 1
            N, C, H, W = x.shape
            win = x.reshape(N, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(N, C, H // 2, W // 2, 4)
            out[l] = torch.argmax((win == win.max(dim=4, keepdim=True).values).to(torch.uint8), dim=4).permute(0, 2, 3, 1).to(torch.uint8).numpy()
    return out


def features(images_nhwc, weights):
    return forward(images_nhwc, weights)[1]


def relu_masks(images_nhwc, weights):
    """The oracle's own ReLU decisions {conv layer l < 12: bool [N, H, W, C]} = [z_l > 0] (same format as Forced.masks)."""
    xs, _ = forward(images_nhwc, weights)
    out, j = {}, 0
    for l in range(len(weights) - 1):
        j += 1                      # xs[j] is relu(z_l): the pool's input when a pool follows, else the next conv's input
        out[l] = (xs[j] > 0).permute(0, 2, 3, 1).numpy()
        if l in _pools(weights):
            j += 1
    return out


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


def analyze(method, images_nhwc, R_head_nhwc, weights, epsilon=1e-7, alpha=None, beta=None, bias=True, force=None):
    """Relevance / gradient of the supplied head tensor at block5_conv3's output, at the input image.

    method: 'lrp.epsilon' | 'lrp.z' | 'lrp.alpha_beta' | 'lrp.alpha_1_beta_0' | 'lrp.alpha_2_beta_1' |
            'lrp.z_plus' | 'lrp.z_plus_fast' | 'lrp.sequential_preset_a' |
            'gradient' | 'input_t_gradient' | 'guided_backprop'
    """
    xs, _ = forward(images_nhwc, weights, force=force)
    R = _to_nchw(R_head_nhwc)
    img = xs[0]
    if method in ("gradient", "input_t_gradient", "guided_backprop"):
        j = len(xs) - 1
        for l in range(len(weights) - 1, -1, -1):
            k, b = weights[l]
            if l in _pools(weights):
                R = _grad(lambda v: _pool(v, l, force), xs[j], R)
                j -= 1
            if method == "guided_backprop":
                R = torch.relu(R)
            R = _grad(lambda v: _relu(_conv(v, k, b), l, force), xs[j], R)
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
        if l in _pools(weights):
            R = _grad(lambda v: _pool(v, l, force), xs[j], R)
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
