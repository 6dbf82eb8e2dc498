"""Host-side handle of the CUDA attention-LSTM decoder (include/lrpcap.h, "decoder" section)."""
import ctypes

import numpy as np
import torch

from . import _lib

_SHARED = ("image_features_w", "image_features_b", "global_w", "global_b", "embedding", "output_w", "output_b")
_ADAPTIVE = ("lstm_wi", "lstm_wh", "lstm_b", "Wv", "Wg", "Wx", "Wh", "Ws")
_GRIDTD = ("lang_wi", "lang_wh", "lang_b", "td_wi", "td_wh", "td_b", "W_va", "W_ha", "W_a", "W_x", "W_h", "W_s")


class DecoderEngine(object):
    """dec: weight dict as produced by synth.decoder_weights / model.CaptioningModel (Keras layouts, float32)."""

    def __init__(self, dec, sos=1, keras_logits=False, device="cuda:0"):
        self.kind = dec["kind"]
        self.H, self.E, self.D, self.V = dec["hidden_dim"], dec["embedding_dim"], dec["D"], dec["vocab_size"]
        self.sos = int(sos)
        self.keras_logits = bool(keras_logits)
        self.device = torch.device(device)
        names = _SHARED + (_ADAPTIVE if self.kind == "adaptive" else _GRIDTD)
        self._w = {k: np.ascontiguousarray(dec[k], dtype=np.float32) for k in names}
        if self.kind == "adaptive":
            self._w["Vatt"] = np.ascontiguousarray(dec["V"], dtype=np.float32)   # attention vector `_V` (H, 1)
        self._h = None
        self.N = self.T = self.L = 0

    def handle(self):
        if self._h is None:
            if not torch.cuda.is_available():
                raise _lib.LrpcapError(-3, "no CUDA device: lrpcap has no CPU path")
            lib = _lib.load()
            torch.cuda.set_device(self.device)
            w = _lib.DecoderWeights()
            w.kind = _lib.DECODER_ADAPTIVE if self.kind == "adaptive" else _lib.DECODER_GRIDTD
            w.V, w.H, w.E, w.D = self.V, self.H, self.E, self.D
            for k, v in self._w.items():
                setattr(w, k, _lib.fptr(v))
            h = _lib.c_void_p()
            _lib.check(lib.lrpcap_decoder_create(ctypes.byref(h), ctypes.byref(w), self.sos, int(self.keras_logits)))
            self._h = h
        return self._h

    def set_weights_device(self, tensors):
        """In-place weight replacement from CUDA float32 tensors: `tensors` maps the weight names of synth.decoder_weights
        ('image_features_w', ..., 'V' for the attention vector) to contiguous tensors of the handle's shapes."""
        names = _SHARED + (_ADAPTIVE if self.kind == "adaptive" else _GRIDTD)
        w = _lib.DecoderWeights()
        w.kind = _lib.DECODER_ADAPTIVE if self.kind == "adaptive" else _lib.DECODER_GRIDTD
        w.V, w.H, w.E, w.D = self.V, self.H, self.E, self.D
        keep = []
        for k in names + (("Vatt",) if self.kind == "adaptive" else ()):
            t = tensors["V" if k == "Vatt" else k]
            ref = self._w[k]
            if tuple(t.shape) != ref.shape or t.dtype != torch.float32 or not t.is_cuda:
                raise ValueError("set_weights_device: %s must be a float32 CUDA tensor of shape %s" % (k, ref.shape))
            t = t.contiguous()
            keep.append(t)
            setattr(w, k, ctypes.cast(t.data_ptr(), _lib.c_float_p))
        _lib.check(_lib.load().lrpcap_decoder_set_weights_device(self.handle(), ctypes.byref(w)))
        self.N = 0

    def predict(self, features, captions, eos=-1):
        """Teacher-forced forward that returns the arg-max token of every step ([N, T] int32) -- `argmax(model.predict)` of
        the LRP-inference loop.  `captions[n, i]` is the token fed at step i + 1 (SOS is fed at step 0).  The handle holds
        no explainable state afterwards: run `forward(features, captions=predicted)` next."""
        f = features
        if isinstance(f, np.ndarray):
            f = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32))
        f = f.to(self.device).contiguous().float()
        N = f.shape[0]
        f = f.reshape(N, -1, f.shape[-1])
        cap = np.ascontiguousarray(np.asarray(captions), dtype=np.int32).copy()
        if cap.ndim != 2 or cap.shape[0] != N:
            raise ValueError("captions must be [N, T]")
        _lib.check(_lib.load().lrpcap_decoder_forward(self.handle(), _lib.c_void_p(f.data_ptr()), N, f.shape[1], _lib.iptr(cap),
                                                      cap.shape[1], 2, int(eos), self._stream()))
        self.N = 0
        return cap

    def close(self):
        if self._h is not None:
            _lib.load().lrpcap_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return _lib.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self, features, captions=None, T=None, greedy=False, eos=-1):
        """features: [N, L, D] or [N, h, w, D] (torch cuda / numpy). captions: int [N, T] tokenizer ids
        (ignored and returned when greedy). Returns the captions as int32 numpy [N, T]."""
        f = features
        if isinstance(f, np.ndarray):
            f = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32))
        f = f.to(self.device).contiguous().float()
        N = f.shape[0]
        f = f.reshape(N, -1, f.shape[-1])
        if f.shape[2] != self.D:
            raise ValueError("feature depth %d != D=%d" % (f.shape[2], self.D))
        L = f.shape[1]
        if greedy:
            if T is None:
                raise ValueError("greedy decoding needs T")
            cap = np.zeros((N, int(T)), dtype=np.int32)
        else:
            cap = np.ascontiguousarray(np.asarray(captions), dtype=np.int32)
            if cap.ndim != 2 or cap.shape[0] != N:
                raise ValueError("captions must be [N, T]")
        _lib.check(_lib.load().lrpcap_decoder_forward(self.handle(), _lib.c_void_p(f.data_ptr()), N, L, _lib.iptr(cap),
                                                      cap.shape[1], int(bool(greedy)), int(eos), self._stream()))
        self._features = f
        self.N, self.T, self.L = N, cap.shape[1], L
        self.captions = cap
        return cap

    def _words(self, word_img, word_t):
        wi = np.ascontiguousarray(np.asarray(word_img), dtype=np.int32)
        wt = np.ascontiguousarray(np.asarray(word_t), dtype=np.int32)
        if wi.shape != wt.shape or wi.ndim != 1:
            raise ValueError("word_img and word_t must be 1-D and equally long")
        return wi, wt

    def relevance(self, word_img, word_t, want_words=True, want_attention=True):
        """LRP of logit(word t of image n) -> (R_head cuda [W, L, D], r_words [W, T] | None, attention [W, L] | None)."""
        wi, wt = self._words(word_img, word_t)
        W = wi.shape[0]
        if np.any(wt > self.T) or np.any(wt < 1):
            raise NotImplementedError("index out of range of captions")
        out = torch.empty((W, self.L, self.D), dtype=torch.float32, device=self.device)
        rw = np.zeros((W, self.T), dtype=np.float64) if want_words else None
        at = np.zeros((W, self.L), dtype=np.float32) if want_attention else None
        _lib.check(_lib.load().lrpcap_decoder_relevance(
            self.handle(), _lib.iptr(wi), _lib.iptr(wt), W, _lib.c_void_p(out.data_ptr()),
            _lib.dptr(rw) if want_words else None, _lib.fptr(at) if want_attention else None, self._stream()))
        return out, rw, at

    def backward(self, word_img, word_t, want_words=True):
        wi, wt = self._words(word_img, word_t)
        W = wi.shape[0]
        out = torch.empty((W, self.L, self.D), dtype=torch.float32, device=self.device)
        rw = np.zeros((W, self.T), dtype=np.float64) if want_words else None
        _lib.check(_lib.load().lrpcap_decoder_backward(
            self.handle(), _lib.iptr(wi), _lib.iptr(wt), W, _lib.c_void_p(out.data_ptr()),
            _lib.dptr(rw) if want_words else None, self._stream()))
        return out, rw

    def caption_logits(self):
        out = np.zeros((self.N, self.T), dtype=np.float64)
        _lib.check(_lib.load().lrpcap_decoder_caption_logits(self.handle(), _lib.dptr(out)))
        return out

    def last_logits(self):
        """Full logits [N, V] of the last step of the most recent forward (float64)."""
        out = np.zeros((self.N, self.V), dtype=np.float64)
        _lib.check(_lib.load().lrpcap_decoder_last_logits(self.handle(), _lib.dptr(out), self._stream()))
        return out

    def attention(self):
        """(alpha [N, T+1, L], beta [N, T+1]) with the zero row at index 0, as the reference stores them."""
        al = np.zeros((self.N, self.T + 1, self.L), dtype=np.float32)
        be = np.zeros((self.N, self.T + 1), dtype=np.float32)
        _lib.check(_lib.load().lrpcap_decoder_attention(self.handle(), _lib.fptr(al), _lib.fptr(be)))
        return al, be

    def launches(self):
        return int(_lib.load().lrpcap_decoder_launches(self.handle()))
