"""Batched two-stage explanation engine: all (image, word) jobs of a batch in single launches.

This is the data-parallel form of the reference's per-word loop
    for each image: _forward_beam_search; for t in words: _explain_lstm_single_word_sequence(t); _explain_CNN(img, R_t)
(explain_image.py:45-87, models/explainers.py:183-189).  Jobs are independent units, so multi-GPU use is a plain
partition of the images over ranks (`shard_images`); the only collective is the optional gather of the maps.
"""
import numpy as np
import torch

from . import _lib
from .decoder import DecoderEngine
from .encoder import RuleSpec

METHOD_LRP, METHOD_GRADIENT = 0, 1


def shard_images(n_images, rank, world_size):
    """Contiguous block partition of image indices; all words of an image stay on one rank."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_images, world_size)
    start = rank * base + min(rank, rem)
    return np.arange(start, start + base + (1 if rank < rem else 0), dtype=np.int64)


def word_list(n_images, T, lengths=None):
    """(word_img, word_t) for every word t = 1..len of every image (row-major: image, then position)."""
    if lengths is None:
        lengths = [T] * n_images
    wi = np.concatenate([np.full(int(l), i, dtype=np.int32) for i, l in enumerate(lengths)]) if n_images else np.zeros(0, np.int32)
    wt = np.concatenate([np.arange(1, int(l) + 1, dtype=np.int32) for l in lengths]) if n_images else np.zeros(0, np.int32)
    return wi, wt


class ExplainEngine(object):
    """model: model.CaptioningModel. rule: encoder RuleSpec (default: the reference's PresetA = alpha1-beta0 with bias)."""

    def __init__(self, model, rule=None, sos=1, eos=2, keras_logits=False):
        self.model = model
        self.image_model = model.image_model
        self.rule = rule if rule is not None else RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True)
        self.sos, self.eos = sos, eos
        self.decoder = DecoderEngine(model.dec, sos=sos, keras_logits=keras_logits, device=model.device)

    def forward(self, images, captions=None, T=None, greedy=False, suppress_eos=True):
        """Encoder + decoder forward. Returns captions [N, T] (generated when greedy)."""
        self.image_model.forward(images, self.rule)
        feats = self.image_model.features()
        return self.decoder.forward(feats, captions=captions, T=T, greedy=greedy,
                                    eos=self.eos if (greedy and suppress_eos) else -1)

    def explain_words(self, word_img, word_t, method=METHOD_LRP, want_side=False):
        """Pixel maps [W, hw, hw, 3] (cuda) for the listed words of the batch given to forward()."""
        if method == METHOD_LRP:
            R_head, rw, att = self.decoder.relevance(word_img, word_t, want_words=want_side, want_attention=want_side)
        else:
            R_head, rw = self.decoder.backward(word_img, word_t, want_words=want_side)
            att = None
        fh = self.image_model.image_hw // 16
        maps = self.image_model.relevance(word_img, R_head.view(-1, fh, fh, R_head.shape[-1]))
        return (maps, R_head, rw, att) if want_side else maps

    def explain_batch(self, images, captions=None, T=None, greedy=False, method=METHOD_LRP):
        cap = self.forward(images, captions=captions, T=T, greedy=greedy)
        wi, wt = word_list(cap.shape[0], cap.shape[1])
        return self.explain_words(wi, wt, method=method), cap

    def explain_batch_host(self, images, captions, greedy=False, method=METHOD_LRP, out=None):
        """One C-ABI call with host buffers in and out (lrpcap_explain_batch_host): images [N, hw, hw, 3] float32,
        captions [N, T] int32 (filled in when greedy); returns maps [N*T, hw, hw, 3] float32 (host)."""
        images = np.ascontiguousarray(images, dtype=np.float32) if not isinstance(images, np.ndarray) or images.dtype != np.float32 else images
        N, hw = images.shape[0], images.shape[1]
        T = captions.shape[1]
        if out is None:
            out = np.empty((N * T, hw, hw, 3), dtype=np.float32)
        r = self.rule
        stream = _lib.c_void_p(torch.cuda.current_stream(self.image_model.device).cuda_stream)
        _lib.check(_lib.load().lrpcap_explain_batch_host(
            self.image_model.handle(), self.decoder.handle(), _lib.fptr(images), N, _lib.iptr(captions), T,
            int(bool(greedy)), int(self.eos if greedy else -1), int(method), r.kind, r.epsilon, r.alpha, r.beta,
            int(r.bias), _lib.fptr(out), stream))
        return out

    def launches(self):
        return self.image_model.launches() + self.decoder.launches()


class StreamedEngine(object):
    """Runs `lanes` independent ExplainEngine instances, one CUDA stream and one host thread each, on contiguous blocks
    of the batch's images.

    Explanation jobs of different images are independent, and the decoder phases are chains of
    small latency-bound kernels that leave most SMs idle: with several lanes in flight those phases overlap each other
    and the tensor-core phases of the other lanes, while the tcgen05 kernels (persistent, one CTA per SM) simply queue
    behind one another. Each lane owns its handles (per-image state, message buffers); the weights are shared on the
    host and uploaded once per lane."""

    def __init__(self, model, rule=None, lanes=2, sos=1, eos=2, keras_logits=False, chunk_words=None):
        import threading
        from .model import CaptioningModel
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        self.device = torch.device(model.device)
        self.lanes = []
        for i in range(lanes):
            m = model if i == 0 else CaptioningModel(model.kind, model.vgg, model.dec, image_hw=model.image_hw,
                                                     precision=model.precision, device=model.device)
            eng = ExplainEngine(m, rule=rule, sos=sos, eos=eos, keras_logits=keras_logits)
            if chunk_words:
                m.image_model.set_chunk_words(chunk_words)
            self.lanes.append({"engine": eng, "stream": torch.cuda.Stream(device=self.device)})
        self.rule = self.lanes[0]["engine"].rule
        self.eos = eos
        self._threading = threading

    def _split(self, n):
        k = len(self.lanes)
        base, rem = divmod(n, k)
        bounds, s = [], 0
        for i in range(k):
            e = s + base + (1 if i < rem else 0)
            bounds.append((s, e))
            s = e
        return bounds

    def _run(self, fn):
        """fn(lane_index, lane) on one host thread per lane; the lane streams start after the caller's stream and the
        caller's stream continues after all of them (so CUDA events on the caller's stream bracket the work)."""
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        results, errors = [None] * len(self.lanes), []

        def work(i, lane):
            try:
                torch.cuda.set_device(self.device)
                lane["stream"].wait_event(start)
                with torch.cuda.stream(lane["stream"]):
                    results[i] = fn(i, lane)
            except BaseException as e:   # re-raised on the calling thread
                errors.append(e)
        threads = [self._threading.Thread(target=work, args=(i, lane)) for i, lane in enumerate(self.lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for lane in self.lanes:
            done = torch.cuda.Event()
            done.record(lane["stream"])
            cur.wait_event(done)
        if errors:
            raise errors[0]
        return results

    def explain_batch(self, images, T, greedy=True, method=METHOD_LRP, captions=None):
        """images: torch cuda [N, hw, hw, 3]. Returns (list of per-lane maps [n_i*T, hw, hw, 3] in image order, captions)."""
        bounds = self._split(images.shape[0])

        def fn(i, lane):
            a, b = bounds[i]
            if b <= a:
                return None, np.zeros((0, T), dtype=np.int32)
            cap = None if captions is None else np.ascontiguousarray(captions[a:b])
            return lane["engine"].explain_batch(images[a:b], captions=cap, T=T, greedy=greedy, method=method)
        res = self._run(fn)
        maps = [r[0] for r in res if r[0] is not None]
        for m in maps:
            m.record_stream(torch.cuda.current_stream(self.device))
        return maps, np.concatenate([r[1] for r in res], axis=0)

    def explain_batch_host(self, images, captions, greedy=True, method=METHOD_LRP, out=None):
        """Host buffers in and out, one lrpcap_explain_batch_host call per lane on its block of images."""
        N, hw, T = images.shape[0], images.shape[1], captions.shape[1]
        if out is None:
            out = np.empty((N * T, hw, hw, 3), dtype=np.float32)
        bounds = self._split(N)

        def fn(i, lane):
            a, b = bounds[i]
            if b > a:
                lane["engine"].explain_batch_host(images[a:b], captions[a:b], greedy=greedy, method=method, out=out[a * T:b * T])
        self._run(fn)
        return out

    def launches(self):
        return sum(l["engine"].launches() for l in self.lanes)
