"""ctypes binding of include/lrpcap.h (liblrpcap.so, built in-tree by build.py).

There is no CPU fallback: if the shared library is missing, or a call fails, this module raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LRPCAP_LIB") or os.path.join(_HERE, "liblrpcap.so")   # LRPCAP_LIB: e.g. the bounds-checking debug build

OK = 0
PREC_FP32_SIMT, PREC_BF16X3_TC, PREC_F16X2_TC, PREC_TC_AUTO, PREC_H1F8_TC = 0, 1, 2, 3, 4
RULE_EPSILON, RULE_Z, RULE_ALPHA_BETA, RULE_ZPLUS_FAST, RULE_GRADIENT, RULE_INPUT_T_GRADIENT, RULE_GUIDED_BACKPROP = range(7)
DECODER_ADAPTIVE, DECODER_GRIDTD = 0, 1

c_float_p = ctypes.POINTER(ctypes.c_float)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_void_p = ctypes.c_void_p


class LrpcapError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("lrpcap error %d: %s" % (code, message))
        self.code = code


class DecoderWeights(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("V", ctypes.c_int), ("H", ctypes.c_int), ("E", ctypes.c_int),
                ("D", ctypes.c_int)] + [(n, c_float_p) for n in (
                    "image_features_w", "image_features_b", "global_w", "global_b", "embedding", "output_w", "output_b",
                    "lstm_wi", "lstm_wh", "lstm_b", "Wv", "Wg", "Wx", "Wh", "Ws", "Vatt",
                    "lang_wi", "lang_wh", "lang_b", "td_wi", "td_wh", "td_b", "W_va", "W_ha", "W_a", "W_x", "W_h", "W_s")]


# name -> (restype, argtypes); must list every symbol include/lrpcap.h declares
PROTOTYPES = {
    "lrpcap_last_error": (ctypes.c_char_p, []),
    "lrpcap_version": (ctypes.c_int, []),
    "lrpcap_encoder_create": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_float_p), ctypes.POINTER(c_float_p), ctypes.c_int, ctypes.c_int]),
    "lrpcap_encoder_create_arch": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.c_int, ctypes.POINTER(c_float_p), ctypes.POINTER(c_float_p), ctypes.c_int, ctypes.c_int]),
    "lrpcap_encoder_set_weights": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_float_p), ctypes.POINTER(c_float_p)]),
    "lrpcap_encoder_set_weights_device": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "lrpcap_encoder_destroy": (ctypes.c_int, [c_void_p]),
    "lrpcap_encoder_forward": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, c_void_p]),
    "lrpcap_encoder_features": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "lrpcap_encoder_relevance": (ctypes.c_int, [c_void_p, c_int_p, c_void_p, ctypes.c_int, c_void_p, c_void_p]),
    "lrpcap_encoder_relevance_host": (ctypes.c_int, [c_void_p, c_int_p, c_float_p, ctypes.c_int, c_float_p, c_void_p]),
    "lrpcap_encoder_set_chunk_words": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "lrpcap_encoder_set_promote": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "lrpcap_encoder_launches": (ctypes.c_longlong, [c_void_p]),
    "lrpcap_encoder_profile": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "lrpcap_encoder_profile_read": (ctypes.c_int, [c_void_p, c_double_p]),
    "lrpcap_decoder_create": (ctypes.c_int, [ctypes.POINTER(c_void_p), ctypes.POINTER(DecoderWeights), ctypes.c_int, ctypes.c_int]),
    "lrpcap_decoder_set_weights_device": (ctypes.c_int, [c_void_p, ctypes.POINTER(DecoderWeights)]),
    "lrpcap_decoder_destroy": (ctypes.c_int, [c_void_p]),
    "lrpcap_decoder_forward": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p]),
    "lrpcap_decoder_relevance": (ctypes.c_int, [c_void_p, c_int_p, c_int_p, ctypes.c_int, c_void_p, c_double_p, c_float_p, c_void_p]),
    "lrpcap_decoder_backward": (ctypes.c_int, [c_void_p, c_int_p, c_int_p, ctypes.c_int, c_void_p, c_double_p, c_void_p]),
    "lrpcap_decoder_caption_logits": (ctypes.c_int, [c_void_p, c_double_p]),
    "lrpcap_decoder_attention": (ctypes.c_int, [c_void_p, c_float_p, c_float_p]),
    "lrpcap_decoder_last_logits": (ctypes.c_int, [c_void_p, c_double_p, c_void_p]),
    "lrpcap_decoder_launches": (ctypes.c_longlong, [c_void_p]),
    "lrpcap_explain_batch_host": (ctypes.c_int, [c_void_p, c_void_p, c_float_p, ctypes.c_int, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, c_float_p, c_void_p]),
    "lrpcap_gradcam": (ctypes.c_int, [c_void_p, c_int_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, c_void_p, c_void_p]),
    "lrpcap_scale_maps": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_void_p]),
    "lrpcap_lrp_inference_scores": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, c_void_p]),
    "lrpcap_heatmaps": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_float_p, c_void_p]),
    "lrpcap_bbox_correctness": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int, c_float_p, ctypes.c_int, c_float_p, c_void_p]),
    "lrpcap_encoder_debug_pool_routes": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p]),
    "lrpcap_encoder_debug_multiplier": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, c_float_p]),
    "lrpcap_encoder_debug_message_scales": (ctypes.c_int, [c_void_p, c_float_p, c_int_p, ctypes.c_int, c_int_p]),
    "lrpcap_debug_conv_tile": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "lrpcap_debug_conv": (ctypes.c_int, [ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_int, c_float_p]),
}

_lib = None


def load():
    """Loads liblrpcap.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LrpcapError(-5, "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(the CUDA library is the only implementation)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code):
    if code != OK:
        raise LrpcapError(code, load().lrpcap_last_error().decode("utf-8", "replace"))


def fptr(a):
    """Host float32 C-contiguous array -> float*."""
    assert isinstance(a, np.ndarray) and a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_float_p)


def iptr(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


def dptr(a):
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def debug_conv_tile(items, H, W):
    """(tile_w, tile_h, tile_items) of the generic tcgen05 kernel for an [items, H, W] map (host-only call)."""
    lib = load()
    tw, th, ti = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib.lrpcap_debug_conv_tile(int(items), int(H), int(W), ctypes.byref(tw), ctypes.byref(th), ctypes.byref(ti)))
    return tw.value, th.value, ti.value


def debug_conv(precision, A, B, taps):
    """A [items,H,W,C] f32, B [taps,C,Nout] f32 -> [items,H,W,Nout] f32 through one GEMM kernel."""
    lib = load()
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    items, H, W, C = A.shape
    Nout = B.shape[-1]
    out = np.empty((items, H, W, Nout), dtype=np.float32)
    check(lib.lrpcap_debug_conv(precision, fptr(A), items, H, W, C, fptr(B), taps, Nout, fptr(out)))
    return out
