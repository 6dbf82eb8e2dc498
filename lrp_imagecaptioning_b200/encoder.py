"""Host-side handle of the CUDA VGG16 encoder (include/lrpcap.h, "encoder" section).

`ImageModel` plays the role of the reference's `_image_model` (Keras sub-model input_1 -> block5_conv3,
models/explainers.py:29-30): it owns the 13 conv kernels/biases with their Keras names and `predict`s grid
features.  The relevance rules are chosen per call (analyzers.py).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .synth import VGG16_CFG, VGG19_CFG


class RuleSpec(object):
    def __init__(self, kind, epsilon=1e-7, alpha=1.0, beta=0.0, bias=True):
        self.kind, self.epsilon, self.alpha, self.beta, self.bias = kind, float(epsilon), float(alpha), float(beta), bool(bias)

    def key(self):
        return (self.kind, self.epsilon, self.alpha, self.beta, self.bias)


def _as_cuda_f32(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if not x.is_cuda:
        x = x.to(device, non_blocking=True)
    return x.contiguous().float()


class ImageModel(object):
    """VGG16 (or VGG19) conv stack with Keras layer names; `weights` = list of 13 (16) (kernel HWIO, bias)."""
    layer_names = [c[0] for c in VGG16_CFG]

    def __init__(self, weights, image_hw=224, precision="bf16x3", device="cuda:0"):
        if len(weights) not in (13, 16):
            raise ValueError("VGG16 to block5_conv3 has 13 conv layers (VGG19 to block5_conv4: 16), got %d" % len(weights))
        self.cfg = VGG16_CFG if len(weights) == 13 else VGG19_CFG
        self.arch = 0 if len(weights) == 13 else 1
        self.layer_names = [c[0] for c in self.cfg]
        for (k, b), (name, cin, cout, _) in zip(weights, self.cfg):
            if tuple(k.shape) != (3, 3, cin, cout) or tuple(b.shape) != (cout,):
                raise ValueError("layer %s: expected kernel (3,3,%d,%d) and bias (%d,)" % (name, cin, cout, cout))
        self.weights = [(np.ascontiguousarray(k, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32))
                        for k, b in weights]
        self.image_hw = int(image_hw)
        self.precision = {"fp32": _lib.PREC_FP32_SIMT, "bf16x3": _lib.PREC_BF16X3_TC, "f16x2": _lib.PREC_F16X2_TC,
                          "tc": _lib.PREC_TC_AUTO, "h1f8": _lib.PREC_H1F8_TC}[precision]
        self.device = torch.device(device)
        self._h = None
        self._state = None   # (rule key, n_images) of the resident forward state

    # -- handle management
    def handle(self):
        if self._h is None:
            if not torch.cuda.is_available():
                raise _lib.LrpcapError(-3, "no CUDA device: lrpcap has no CPU path")
            lib = _lib.load()
            torch.cuda.set_device(self.device)
            n = len(self.weights)
            ks = (_lib.c_float_p * n)(*[_lib.fptr(k) for k, _ in self.weights])
            bs = (_lib.c_float_p * n)(*[_lib.fptr(b) for _, b in self.weights])
            h = _lib.c_void_p()
            _lib.check(lib.lrpcap_encoder_create_arch(ctypes.byref(h), self.arch, ks, bs, self.image_hw, self.precision))
            self._h = h
        return self._h

    def set_weights(self, weights):
        """Replace the 13 (kernel, bias) pairs in place (keeps the handle's large buffers; the forward state is dropped)."""
        weights = [(np.ascontiguousarray(k, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)) for k, b in weights]
        if len(weights) != len(self.weights) or any(k.shape != k0.shape or b.shape != b0.shape for (k, b), (k0, b0) in zip(weights, self.weights)):
            raise ValueError("set_weights needs %d (kernel, bias) pairs with the shapes the model was built with" % len(self.weights))
        self.weights = weights
        self._state = None
        if self._h is not None:
            n = len(self.weights)
            ks = (_lib.c_float_p * n)(*[_lib.fptr(k) for k, _ in self.weights])
            bs = (_lib.c_float_p * n)(*[_lib.fptr(b) for _, b in self.weights])
            _lib.check(_lib.load().lrpcap_encoder_set_weights(self._h, ks, bs))

    def set_weights_device(self, kernels_hwio, biases):
        """In-place weight replacement from CUDA tensors (lists of contiguous float32 tensors, HWIO kernels): no host round
        trip.  `self.weights` (the host copy) is NOT refreshed -- call `sync_host_weights` if it is needed."""
        if len(kernels_hwio) != len(self.weights) or len(biases) != len(self.weights):
            raise ValueError("set_weights_device needs %d kernels and biases" % len(self.weights))
        for k, b, (k0, b0) in zip(kernels_hwio, biases, self.weights):
            if tuple(k.shape) != k0.shape or tuple(b.shape) != b0.shape or k.dtype != torch.float32 or not k.is_cuda or not k.is_contiguous():
                raise ValueError("set_weights_device: contiguous float32 CUDA tensors with the model's shapes are required")
        n = len(self.weights)
        ks = (_lib.c_void_p * n)(*[k.data_ptr() for k in kernels_hwio])
        bs = (_lib.c_void_p * n)(*[b.contiguous().data_ptr() for b in biases])
        self._state = None
        _lib.check(_lib.load().lrpcap_encoder_set_weights_device(self.handle(), ks, bs))

    def close(self):
        if self._h is not None:
            _lib.load().lrpcap_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return _lib.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- compute
    def forward(self, images, rule):
        """images: [N, hw, hw, 3] float32 (numpy or torch); builds features + the rule's per-image state."""
        self.handle()   # fails loudly first when there is no CUDA device / library
        x = _as_cuda_f32(images, self.device)
        if x.dim() != 4 or x.shape[1] != self.image_hw or x.shape[2] != self.image_hw or x.shape[3] != 3:
            raise ValueError("images must be [N, %d, %d, 3], got %s" % (self.image_hw, self.image_hw, tuple(x.shape)))
        lib = _lib.load()
        _lib.check(lib.lrpcap_encoder_forward(self.handle(), _lib.c_void_p(x.data_ptr()), x.shape[0], rule.kind,
                                              rule.epsilon, rule.alpha, rule.beta, int(rule.bias), self._stream()))
        self._state = (rule.key(), x.shape[0])
        self._images = x   # keep alive until the async copy inside forward has been consumed
        return self

    def features(self):
        n = self._state[1]
        fh = self.image_hw // 16
        out = torch.empty((n, fh, fh, 512), dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().lrpcap_encoder_features(self.handle(), _lib.c_void_p(out.data_ptr()), self._stream()))
        return out

    def predict(self, images):
        """Keras-style: grid features as a numpy array [N, hw/16, hw/16, 512]."""
        self.forward(images, RuleSpec(_lib.RULE_GRADIENT))
        return self.features().cpu().numpy()

    def relevance(self, img_index, R_head):
        """img_index: int array [W]; R_head [W, fh, fh, 512] (torch cuda or numpy). Returns torch cuda [W, hw, hw, 3]."""
        idx = np.ascontiguousarray(np.asarray(img_index), dtype=np.int32)
        R = _as_cuda_f32(R_head, self.device)
        fh = self.image_hw // 16
        if tuple(R.shape) != (idx.shape[0], fh, fh, 512):
            raise ValueError("R_head must be [%d, %d, %d, 512], got %s" % (idx.shape[0], fh, fh, tuple(R.shape)))
        out = torch.empty((idx.shape[0], self.image_hw, self.image_hw, 3), dtype=torch.float32, device=self.device)
        _lib.check(_lib.load().lrpcap_encoder_relevance(self.handle(), _lib.iptr(idx), _lib.c_void_p(R.data_ptr()),
                                                        idx.shape[0], _lib.c_void_p(out.data_ptr()), self._stream()))
        return out

    def set_chunk_words(self, n):
        _lib.check(_lib.load().lrpcap_encoder_set_chunk_words(self.handle(), int(n)))

    def set_promote(self, every_k_steps):
        _lib.check(_lib.load().lrpcap_encoder_set_promote(self.handle(), int(every_k_steps)))

    def profile(self, enable=True):
        _lib.check(_lib.load().lrpcap_encoder_profile(self.handle(), int(bool(enable))))

    def profile_read(self):
        """{class: (ms, algorithmic FLOPs, launches)} for classes tc_bwd / tc_fwd / simt / last; resets the counters."""
        out = np.zeros(12, dtype=np.float64)
        _lib.check(_lib.load().lrpcap_encoder_profile_read(self.handle(), _lib.dptr(out)))
        return {k: tuple(out[3 * i:3 * i + 3]) for i, k in enumerate(("tc_bwd", "tc_fwd", "simt", "last"))}

    # -- diagnostics for the parity tests (discrete decisions of the resident forward state)
    def pool_routes(self):
        """{conv layer index: uint8 [N, H/2, W/2, C]} window position (sy*2+sx) each pooled element routes to."""
        n = self._state[1]
        out = {}
        for l, (_, _, cout, pool) in enumerate(self.cfg):
            if not pool:
                continue
            ho = (self.image_hw >> sum(1 for c in self.cfg[:l] if c[3])) // 2
            a = np.empty((n, ho, ho, cout), dtype=np.uint8)
            _lib.check(_lib.load().lrpcap_encoder_debug_pool_routes(self.handle(), l, _lib.c_void_p(a.ctypes.data)))
            out[l] = a
        return out

    def multiplier(self, layer, branch=0):
        """Dense per-image multiplier G of conv layer `layer` (< 12), [N, H, W, C] float32, pool routing folded in."""
        n = self._state[1]
        h = self.image_hw >> sum(1 for c in self.cfg[:layer] if c[3])
        a = np.empty((n, h, h, self.cfg[layer][2]), dtype=np.float32)
        _lib.check(_lib.load().lrpcap_encoder_debug_multiplier(self.handle(), int(layer), int(branch), _lib.fptr(a)))
        return a

    def message_scales(self, cap_words=4096):
        """Two-product backward: (max [layers+1, chunk] float32, log2 scale [layers, chunk] int32) of the last chunk."""
        n = len(self.weights)
        mx = np.zeros((n + 1) * cap_words, dtype=np.float32)
        kt = np.zeros(n * cap_words, dtype=np.int32)
        chunk = ctypes.c_int(0)
        _lib.check(_lib.load().lrpcap_encoder_debug_message_scales(self.handle(), _lib.fptr(mx), _lib.iptr(kt), cap_words,
                                                                   ctypes.byref(chunk)))
        c = chunk.value
        return mx[:(n + 1) * c].reshape(n + 1, c), kt[:n * c].reshape(n, c)

    def launches(self):
        return int(_lib.load().lrpcap_encoder_launches(self.handle()))
