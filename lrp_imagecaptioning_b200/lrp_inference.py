"""LRP-inference: per-word loss weights from pixel-level explanations, and the fine-tuning step that uses them.

Mirrors of the reference:
  LRPInferenceLayerAdaptive / LRPInferenceLayergridTD   models/model.py:1379-1691, :1693-2062   (`call([captions, imgs, y_pred])`)
  Img*LRPInferenceModel (two outputs, loss 0.5/0.5)      models/model.py:1252-1374
  fine-tune loop body                                    train.py:569-577, :648-657

The explanation part (all non-stop-words of all samples) runs through the CUDA engine in one batched pass; the
differentiable model pass (teacher-forced logits, cross-entropy, Adam with clipvalue) is plain PyTorch autograd, which
SURVEY.md section 7 allows for this step; gradients are all-reduced over NCCL when torch.distributed is initialised.
No gradient flows through the LRP weights (models/model.py:1252-1253).
"""
import numpy as np
import torch
import torch.nn.functional as Fn

from . import _lib
from .encoder import RuleSpec
from .engine import ExplainEngine
from .model import CaptioningModel

# nltk.corpus.stopwords.words('english') is not available offline; callers may pass their own list.
DEFAULT_STOP_WORDS = frozenset("""i me my myself we our ours ourselves you your yours yourself yourselves he him his himself she her
hers herself it its itself they them their theirs themselves what which who whom this that these those am is are was were be been
being have has had having do does did doing a an the and but if or because as until while of at by for with about against between
into through during before after above below to from up down in out on off over under again further then once here there when where
why how all any both each few more most other some such no nor not only own same so than too very s t can will just don should now
""".split())


class _LRPInferenceLayer(object):
    _kind = None

    def __init__(self, model, dataset_provider, hidden_dim, embedding_dim, L, D, img_encoder, lrp_inference_mode,
                 stop_words=DEFAULT_STOP_WORDS, overflow="raise"):
        # overflow: the reference writes the weight at column `tokenizer id` of a (V,) buffer (quirk B10), so a predicted
        # id == V raises IndexError there; "raise" reproduces that, "skip" leaves such a word unweighted (training loops)
        self._overflow = overflow
        if model.kind != self._kind:
            raise ValueError("%s needs a %r model" % (type(self).__name__, self._kind))
        if lrp_inference_mode not in ("mean", "pos_mean", "quantile"):
            raise NotImplementedError("the lrp inference mode is not available")
        self._preprocessor = dataset_provider.caption_preprocessor
        self._EOS_ENCODED = self._preprocessor.EOS_TOKEN_LABEL_ENCODED
        self._SOS_ENCODER = self._preprocessor.SOS_TOKEN_LABEL_ENCODED
        self._dataset_provider = dataset_provider
        self._max_caption_length = 20
        self._hidden_dim, self._embedding_dim, self.L, self.D = hidden_dim, embedding_dim, L, D
        self._img_encoder = img_encoder
        self._color_conversion = "BGRtoRGB"
        self._lrp_inference_mode = lrp_inference_mode
        self._stop_words = frozenset(stop_words)
        self._model = model
        # reference: LRPSequentialPresetA(image_model, epsilon=EPS, 'replace') == alpha1-beta0 with bias on VGG16
        self._engine = ExplainEngine(model, rule=RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True),
                                     sos=self._SOS_ENCODER, eos=self._EOS_ENCODED)

    def _word(self, token):
        w = getattr(self._preprocessor, "_word_of", None)
        return w.get(int(token)) if w is not None else None

    def word_list(self, caption_encoded):
        """(sample, position) of the words that get a weight: stop-words skipped, stop at EOS (model.py:1664-1671)."""
        wi, wt = [], []
        for b, cap in enumerate(caption_encoded):
            for i, tok in enumerate(cap):
                if self._word(tok) in self._stop_words:
                    continue
                if tok == self._EOS_ENCODED:
                    break
                wi.append(b)
                wt.append(i + 1)
        return np.asarray(wi, dtype=np.int32), np.asarray(wt, dtype=np.int32)

    def scores(self, maps):
        """maps [W, hw, hw, 3] cuda -> per-word score (float64 numpy)."""
        W, hw = maps.shape[0], maps.shape[1]
        if self._lrp_inference_mode == "quantile":
            hp = maps.mean(dim=-1).reshape(W, -1)
            amax = hp.abs().amax(dim=1, keepdim=True)
            hp = torch.where(amax > 0, hp / amax.clamp_min(1e-38), torch.zeros_like(hp))
            return torch.quantile(hp.double(), 0.9, dim=1).cpu().numpy()
        out = np.zeros(W, dtype=np.float32)
        stream = _lib.c_void_p(torch.cuda.current_stream(maps.device).cuda_stream)
        _lib.check(_lib.load().lrpcap_lrp_inference_scores(_lib.c_void_p(maps.data_ptr()), W, hw,
                                                          0 if self._lrp_inference_mode == "mean" else 1,
                                                          _lib.fptr(out), stream))
        return out.astype(np.float64)

    def sparse_weights(self, images, caption_encoded, features_ready=False):
        """The non-trivial entries of `call`'s result for an already known predicted caption [B, T] (tokenizer ids):
        (sample, position - 1, vocabulary column, score) arrays -- everything else of the (B, T, V) weight is 1.
        features_ready: the engine's encoder already holds the forward state of `images` (same rule)."""
        caption_encoded = np.ascontiguousarray(caption_encoded, dtype=np.int32)
        wi, wt = self.word_list(caption_encoded)
        if not len(wi):
            z = np.zeros(0, dtype=np.int64)
            return z, z, z, np.zeros(0, dtype=np.float64)
        eng = self._engine
        if not features_ready:
            eng.image_model.forward(images, eng.rule)
        eng.decoder.forward(eng.image_model.features(), captions=caption_encoded)
        sc = self.scores(eng.explain_words(wi, wt))
        col = caption_encoded[wi, wt - 1].astype(np.int64)          # quirk B10: column = tokenizer id, not id - 1
        return wi.astype(np.int64), (wt - 1).astype(np.int64), col, sc

    def call(self, inputs):
        assert len(inputs) == 3
        _, img_inputs, y_preds = inputs
        y_preds = np.asarray(y_preds)
        B, T, V = y_preds.shape
        caption_encoded = (np.argmax(y_preds, axis=-1) + 1).astype(np.int32)
        out = np.zeros(y_preds.shape)
        wi, wt = self.word_list(caption_encoded)
        if len(wi):
            self._engine.forward(img_inputs, captions=caption_encoded)
            maps = self._engine.explain_words(wi, wt)
            sc = self.scores(maps)
            for b, t, s in zip(wi, wt, sc):
                # quirk B10: the weight goes to column `word_encode` (tokenizer id), not id-1; id == V overflows as in numpy
                if caption_encoded[b, t - 1] >= V and self._overflow == "skip":
                    continue
                out[b, t - 1, caption_encoded[b, t - 1]] = s
        return 1 + out


class LRPInferenceLayerAdaptive(_LRPInferenceLayer):
    _kind = "adaptive"


class LRPInferenceLayergridTD(_LRPInferenceLayer):
    _kind = "gridtd"


# ------------------------------------------------------------------------------------------- differentiable model
class CaptionerTorch(torch.nn.Module):
    """Teacher-forced captioner with the Keras step math (models/model.py:573-604 adaptive, :668-682, :784-823 grid-TD;
    heads :444-468, :630-660), parameters initialised from / exported to a CaptioningModel."""

    def __init__(self, model):
        super().__init__()
        self.kind, self.hw = model.kind, model.image_hw
        self.H, self.E = model._hidden_dim, model._embedding_dim
        dev = torch.device(model.device)
        self.conv_w = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(k).permute(3, 2, 0, 1).contiguous().to(dev)) for k, _ in model.vgg])
        self.conv_b = torch.nn.ParameterList([torch.nn.Parameter(torch.from_numpy(b).clone().to(dev)) for _, b in model.vgg])
        self.dec = torch.nn.ParameterDict({k: torch.nn.Parameter(torch.from_numpy(np.asarray(v)).clone().to(dev))
                                           for k, v in model.dec.items() if isinstance(v, np.ndarray)})

    def export(self, model):
        """New CaptioningModel with the current parameters (the CUDA engine copies weights at handle creation)."""
        vgg = [(w.detach().permute(2, 3, 1, 0).contiguous().cpu().numpy(), b.detach().cpu().numpy()) for w, b in zip(self.conv_w, self.conv_b)]
        dec = dict(model.dec)
        for k, p in self.dec.items():
            dec[k] = p.detach().cpu().numpy()
        return CaptioningModel(model.kind, vgg, dec, image_hw=model.image_hw, precision=model.precision, device=model.device)

    def export_into(self, model):
        """Writes the current parameters into `model` in place: the encoder handle keeps its buffers
        (lrpcap_encoder_set_weights); decoder engines built from `model.dec` afterwards see the new weights."""
        vgg = [(w.detach().permute(2, 3, 1, 0).contiguous().cpu().numpy(), b.detach().cpu().numpy()) for w, b in zip(self.conv_w, self.conv_b)]
        for k, p in self.dec.items():
            model.dec[k] = p.detach().cpu().numpy()
        model.vgg = vgg
        model.image_model.set_weights(vgg)
        return model

    def sync_engine(self, engine):
        """Device-to-device: the engine's encoder and decoder handles take the current parameters in place
        (lrpcap_encoder_set_weights_device / lrpcap_decoder_set_weights_device); nothing goes through the host.
        The host copies `model.vgg` / `model.dec` are left as they were (use export_into for those)."""
        with torch.no_grad():
            ks = [w.detach().permute(2, 3, 1, 0).contiguous() for w in self.conv_w]
            bs = [b.detach().contiguous() for b in self.conv_b]
            engine.image_model.set_weights_device(ks, bs)
            engine.decoder.set_weights_device({k: p.detach() for k, p in self.dec.items()})

    def features(self, images_nhwc):
        x = images_nhwc.permute(0, 3, 1, 2)
        for l, (w, b) in enumerate(zip(self.conv_w, self.conv_b)):
            x = torch.relu(Fn.conv2d(x, w, b, padding=1))
            if l in (1, 3, 6, 9):
                x = Fn.max_pool2d(x, 2, 2)
        return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1, x.shape[1])

    @staticmethod
    def _lstm(x, h, c, wi, wh, b, H):
        z = x @ wi + h @ wh + b
        i, f, g, o = torch.sigmoid(z[:, :H]), torch.sigmoid(z[:, H:2 * H]), torch.tanh(z[:, 2 * H:3 * H]), torch.sigmoid(z[:, 3 * H:])
        c = f * c + i * g
        return o * torch.tanh(c), c

    def forward(self, tokens_in, images_nhwc):
        """tokens_in [B, T] tokenizer ids fed at each step (SOS first); returns logits [B, T, V]."""
        d, H = self.dec, self.H
        F = self.features(images_nhwc)
        Vf = torch.relu(F @ d["image_features_w"] + d["image_features_b"])
        g = torch.relu(F.mean(dim=1) @ d["global_w"] + d["global_b"])
        B, T = tokens_in.shape
        emb = d["embedding"][tokens_in - 1]
        z = torch.zeros(B, H, device=F.device)
        outs = []
        if self.kind == "adaptive":
            P = Vf @ d["Wv"]
            h, c = z, z
            for t in range(T):
                x = torch.cat([emb[:, t], g], dim=1)
                hn, c = self._lstm(x, h, c, d["lstm_wi"], d["lstm_wh"], d["lstm_b"], H)
                s = torch.tanh(c) * torch.sigmoid(x @ d["Wx"] + h @ d["Wh"])
                hp = hn @ d["Wg"]
                e = torch.tanh(P + hp[:, None]) @ d["V"]
                zs = torch.tanh(s @ d["Ws"] + hp) @ d["V"]
                alpha = torch.softmax(e, dim=1)
                beta = torch.softmax(torch.cat([e, zs[:, None]], dim=1), dim=1)[:, -1]
                chat = beta * s + (1 - beta) * (alpha * Vf).sum(dim=1)
                outs.append((hn + chat) @ d["output_w"] + d["output_b"])
                h = hn
        else:
            P = Vf @ d["W_va"]
            h1, c1, h2, c2 = z, z, z, z
            for t in range(T):
                x1 = torch.cat([h2, g, emb[:, t]], dim=1)
                h1n, c1 = self._lstm(x1, h1, c1, d["td_wi"], d["td_wh"], d["td_b"], H)
                s = torch.tanh(c1) * torch.sigmoid(x1 @ d["W_x"] + h1 @ d["W_h"])
                hp = h1n @ d["W_ha"]
                e = torch.tanh(P + hp[:, None]) @ d["W_a"]
                zs = torch.tanh(s @ d["W_s"] + hp) @ d["W_a"]
                alpha = torch.softmax(e, dim=1)
                beta = torch.softmax(torch.cat([e, zs[:, None]], dim=1), dim=1)[:, -1]
                chat = beta * s + (1 - beta) * (alpha * Vf).sum(dim=1)
                h2, c2 = self._lstm(torch.cat([chat, h1n], dim=1), h2, c2, d["lang_wi"], d["lang_wh"], d["lang_b"], H)
                outs.append((h2 + chat) @ d["output_w"] + d["output_b"])
                h1 = h1n
        return torch.stack(outs, dim=1)


def lrp_inference_loss(logits, lrp_weight, y_true):
    """0.5 CE(logits) + 0.5 CE(logits * w), last time-step dropped (models/model.py:95-103, 1305-1313)."""
    lg, w, y = logits[:, :-1], lrp_weight[:, :-1], y_true[:, :-1]
    ce = lambda z: -(y * torch.log_softmax(z, dim=-1)).sum(dim=-1).mean()   # noqa: E731
    return 0.5 * ce(lg) + 0.5 * ce(lg * w)


def lrp_inference_loss_sparse(logits, b_idx, t_idx, col, score, target):
    """The same loss with the weight given by its non-trivial entries (w = 1 everywhere else) and the target as class
    indices [B, T] (model index = tokenizer id - 1): no dense (B, T, V) weight / one-hot tensors."""
    weighted = logits.clone()
    if b_idx.numel():
        weighted[b_idx, t_idx, col] = logits[b_idx, t_idx, col] * (1.0 + score)
    V = logits.shape[-1]
    tgt = target[:, :-1].reshape(-1)
    ce = lambda z: Fn.cross_entropy(z[:, :-1].reshape(-1, V), tgt)   # noqa: E731
    return 0.5 * ce(logits) + 0.5 * ce(weighted)


def allreduce_mean_(params, dist_module=None):
    """In place: every parameter's .grad becomes the mean over the ranks of the process group (one flattened all-reduce,
    NCCL over NVLink on the GPU box).  Returns the flattened reduced gradient; no-op (returns the local flat gradient)
    without an initialised multi-rank group."""
    import torch.distributed as dist
    dist = dist_module or dist
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
        off = 0
        for p in params:
            p.grad.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
    return flat


class LRPInferenceTrainer(object):
    """One fine-tuning step = predict -> LRP weights for the predicted words -> train on [y, y] (train.py:569-577).

    Everything stays on the device: the engine's handles follow the trained parameters through
    `CaptionerTorch.sync_engine` (device-to-device), the prediction is the engine's own teacher-forced forward in
    arg-max mode (no second VGG16 pass, no (B, T, V) host copy), the per-word scores enter the loss as a sparse update of
    the logits, and the one collective is the NCCL all-reduce of the flattened gradient."""

    def __init__(self, model, dataset_provider, lrp_inference_mode="mean", learning_rate=1e-4, clipvalue=None,
                 stop_words=DEFAULT_STOP_WORDS, tf32=True):
        self.model = model
        self.provider = dataset_provider
        self.mode = lrp_inference_mode
        self.stop_words = stop_words
        self.net = CaptionerTorch(model)
        self.clipvalue = clipvalue if clipvalue is not None else (0.01 if model.kind == "adaptive" else 0.1)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=learning_rate, eps=1e-7)
        layer_cls = LRPInferenceLayerAdaptive if model.kind == "adaptive" else LRPInferenceLayergridTD
        self.layer = layer_cls(model, dataset_provider, model._hidden_dim, model._embedding_dim, model.L, model.D,
                               model.img_encoder, lrp_inference_mode, stop_words=stop_words, overflow="skip")
        if tf32:   # the differentiable pass is plain PyTorch; TF32 is what TF/Keras use on this class of GPU
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
        self.explained_words = 0          # words explained in the last step
        self.explained_words_total = 0
        self.last_gradient = None         # flattened (all-reduced, unclipped) gradient of the last step, when keep_gradient

    def predict(self, captions_in, images):
        """Predicted caption [B, T] (tokenizer ids) = argmax of the teacher-forced logits + 1, on the engine; leaves the
        encoder's forward state of `images` in place for the explanation."""
        eng = self.layer._engine
        cap_in = np.asarray(captions_in)
        teacher = np.concatenate([cap_in[:, 1:], cap_in[:, -1:]], axis=1).astype(np.int32)   # input of step i+1 = teacher[:, i]
        eng.image_model.forward(images, eng.rule)
        return eng.decoder.predict(eng.image_model.features(), teacher)

    def step(self, captions_in, images, y_true, keep_gradient=False):
        """captions_in [B, T] int (tokenizer ids fed at each step, SOS first), images [B, hw, hw, 3], y_true [B, T, V]
        one-hot or [B, T] class indices (model index = tokenizer id - 1). Returns the loss."""
        dev = next(self.net.parameters()).device
        tok = torch.as_tensor(np.asarray(captions_in), device=dev).long()
        img = images if torch.is_tensor(images) else torch.as_tensor(np.asarray(images, dtype=np.float32))
        img = img.to(dev)
        y = torch.as_tensor(np.asarray(y_true), device=dev)
        target = y.long() if y.dim() == 2 else y.argmax(dim=-1)
        self.net.sync_engine(self.layer._engine)                      # the engine explains the model being trained
        pred = self.predict(captions_in, img)
        b_idx, t_idx, col, sc = self.layer.sparse_weights(img, pred, features_ready=True)
        V = self.model._vocab_size
        keep = col < V                                                # quirk B10 overflow: such a word stays unweighted
        self.explained_words = int(len(b_idx))
        self.explained_words_total += self.explained_words
        tb = lambda a, dt: torch.as_tensor(np.asarray(a)[keep], dtype=dt, device=dev)   # noqa: E731
        self.opt.zero_grad(set_to_none=True)
        loss = lrp_inference_loss_sparse(self.net(tok, img), tb(b_idx, torch.long), tb(t_idx, torch.long), tb(col, torch.long),
                                         tb(sc, torch.float32), target)
        loss.backward()
        params = [p for p in self.net.parameters()]
        flat = allreduce_mean_(params)                # NCCL over NVLink: the one collective of the fine-tune step
        if keep_gradient:
            self.last_gradient = flat
        torch.nn.utils.clip_grad_value_(params, self.clipvalue)
        self.opt.step()
        return float(loss.item())
