"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the decoder stage. Never imported by the product.

A NumPy restatement of the reference's decoder forward pass and its word-level relevance /
gradient back-propagation to the CNN grid features:

  adaptive-attention model   forward   /root/reference/models/explainers.py:370-436
                             LRP       explainers.py:537-666
                             gradient  explainers.py:690-832
  grid-TD model              forward   explainers.py:1092-1178
                             LRP       explainers.py:1180-1321
                             gradient  explainers.py:1344-1532
  linear epsilon-LRP helper            explainers.py:141-144, 156-165

The reference walks one output unit / one grid cell at a time and materialises a dense
(D_in x D_out) attribution matrix per call (identity matrices for element-wise ops).  Here the
same algebra is written in closed form

    lin(R, a, z, W) = a * (W @ (R / stab(z)))        (dense weight, bias_factor = 0)
    ew (R, a, z)    = a * R / stab(z)                 (identity weight)
    stab(z)         = z + (z >= 0 ? +eps : -eps),     eps = keras.backend.epsilon() = 1e-7

with the reference's operand dtypes kept (weights float32 as Keras returns them, float32 stores
into ``r_V`` / ``r_img_feature_input`` / the ``d_*`` gradient buffers), so NumPy's promotion rules
reproduce the reference's float32/float64 mix.  ``faithful=True`` switches the helpers to the
reference's cost profile (dense attribution matrices, per-cell Python loops); it is used only to
time the CPU baseline and to cross-check the closed form.

Pinned (see tests/test_oracle_pinning.py and oracle/make_golden.py): this file is checked against
the reference's own code executed under a Keras import stub, and against fixtures generated that
way and committed under tests/golden/.
"""
import numpy as np
from scipy.special import expit as _sigmoid
from scipy.special import softmax as _softmax

KERAS_EPSILON = 1e-7  # keras.backend.epsilon(); default `eps` of explainers.py:157


def stab(z, eps=KERAS_EPSILON):
    """explainers.py:141-144 -- sign(0) = +1."""
    sgn = np.ones(np.shape(z))
    sgn[np.asarray(z) < 0] = -1
    return z + sgn * eps


class DecoderRef(object):
    def __init__(self, dec, sos=1, eos=2, faithful=False):
        self.kind = dec["kind"]
        self.w = dec
        self.H = dec["hidden_dim"]
        self.E = dec["embedding_dim"]
        self.sos, self.eos = sos, eos
        self.faithful = faithful

    # ------------------------------------------------------------------ helpers
    def _lin(self, r, a, z, W):
        """explainers.py:156-165 with bias_factor=0."""
        r = np.asarray(r).reshape(-1)
        if self.faithful:
            att = np.multiply(W, a[:, None])
            return np.sum(att / stab(z) * r, axis=1)
        return a * np.dot(W, r / stab(z))

    def _ew(self, r, a, z):
        if self.faithful:
            return self._lin(r, a, z, np.identity(len(a)))
        return a * r / stab(z)

    def _lstm(self, x, h, c, Wi, Wh, b):
        """explainers.py:125-139 (and :673-688): Keras gate order i, f, c, o."""
        H = self.H
        z = np.dot(x, Wi)
        z += np.dot(h, Wh)
        z = z + b
        zi, zf, zg, zo = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        ia, fa, ga, oa = _sigmoid(zi), _sigmoid(zf), np.tanh(zg), _sigmoid(zo)
        cn = fa * c + ia * ga
        hn = oa * np.tanh(cn)
        return hn, cn, (zi, zf, zg, zo), (ia, fa, ga, oa)

    # ------------------------------------------------------------------ forward
    def forward(self, F, caption):
        """F: (L, D) float32 grid features; caption: tokenizer ids (model index = id-1)."""
        w, H, E = self.w, self.H, self.E
        F = np.asarray(F, dtype=np.float32)
        self.F = F
        self.L, self.D = F.shape
        self.caption = [int(c) for c in caption]
        # TimeDistributed dense + relu, computed row by row in float32, kept in a float64 buffer
        Vp = np.zeros((self.L, H))
        Vp[:] = np.stack([np.dot(F[l], w["image_features_w"]) + w["image_features_b"] for l in range(self.L)])
        self.Vp, self.Vf = Vp, np.maximum(Vp, 0)
        self.a = np.mean(F, axis=0)
        self.gp = np.dot(self.a, w["global_w"]) + w["global_b"]
        self.g = np.maximum(self.gp, 0)
        if self.kind == "adaptive":
            self._forward_adaptive()
        else:
            self._forward_gridtd()
        return self

    def _emb(self, i):
        tok = self.sos if i == 0 else self.caption[i - 1]
        return self.w["embedding"][tok - 1][None]

    def _forward_adaptive(self):
        w, H, E = self.w, self.H, self.E
        P = np.dot(self.Vf, w["Wv"])
        z32 = np.zeros((1, H), dtype="float32")
        keys = ["h", "c", "zi", "zf", "zg", "zo", "ia", "fa", "ga", "oa", "ctx", "s", "chat"]
        S = {k: [z32] for k in keys}
        S["alpha"] = [np.zeros((1, self.L), dtype="float32")]
        S["beta"] = [np.zeros((1, 1), dtype="float32")]
        xs, logits = [], []
        for i in range(len(self.caption)):
            hm1, cm1 = S["h"][-1], S["c"][-1]
            x = np.hstack((self._emb(i), self.g.reshape(1, E)))
            h, c, zs, acts = self._lstm(x, hm1, cm1, w["lstm_wi"], w["lstm_wh"], w["lstm_b"])
            hp = np.dot(h, w["Wg"])
            e = np.dot(np.tanh(hp + P, dtype="float32"), w["V"])          # (L,1) float32 (explainers.py:413)
            alpha = _softmax(e, axis=0)
            s = np.tanh(c) * _sigmoid(np.dot(x, w["Wx"]) + np.dot(hm1, w["Wh"]))
            zext = np.dot(np.tanh(np.dot(s, w["Ws"]) + hp), w["V"])
            beta = _softmax(np.concatenate((e, zext), axis=0), axis=0)[-1][0]
            ctx = np.sum(alpha * self.Vf, axis=0).reshape(1, H)
            chat = beta * s + (1 - beta) * ctx
            logits.append(np.dot(h + chat, w["output_w"]) + w["output_b"])
            for k, v in zip(keys, [h, c, zs[0], zs[1], zs[2], zs[3], acts[0], acts[1], acts[2], acts[3], ctx, s, chat]):
                S[k].append(v.reshape(1, H))
            S["alpha"].append(alpha.reshape(1, self.L))
            S["beta"].append(np.reshape(beta, (1, 1)))
            xs.append(x)
        self.S = {k: np.vstack(v) for k, v in S.items()}
        self.x = np.vstack(xs)
        self.logits = np.vstack(logits)
        self.attention = self.S["alpha"]

    def _forward_gridtd(self):
        w, H, E = self.w, self.H, self.E
        P = np.dot(self.Vf, w["W_va"])
        z32 = np.zeros((1, H), dtype="float32")
        z64 = np.zeros((1, H))
        keys1 = ["h1", "c1", "zi1", "zf1", "zg1", "zo1", "ia1", "fa1", "ga1", "oa1"]
        keys2 = ["h2", "c2", "zi2", "zf2", "zg2", "zo2", "ia2", "fa2", "ga2", "oa2"]
        S = {k: [z32] for k in keys1 + keys2}
        for k in ("ctx", "s", "chat"):
            S[k] = [z64]
        S["alpha"] = [np.zeros((1, self.L))]
        S["beta"] = [np.zeros((1, 1))]
        x1s, x2s, logits = [], [], []
        for i in range(len(self.caption)):
            h1m, c1m = S["h1"][-1].reshape(1, H), S["c1"][-1].reshape(1, H)
            h2m, c2m = S["h2"][-1].reshape(1, H), S["c2"][-1].reshape(1, H)
            x1 = np.hstack((h2m, self.g.reshape(1, E), self._emb(i)))
            h1, c1, z1, a1 = self._lstm(x1, h1m, c1m, w["td_wi"], w["td_wh"], w["td_b"])
            hp = np.dot(h1, w["W_ha"])
            e = np.dot(np.tanh(P + hp), w["W_a"])
            alpha = _softmax(e, axis=0)
            ctx = np.sum(alpha * self.Vf, axis=0).reshape(1, H)
            s = np.tanh(c1) * _sigmoid(np.dot(x1, w["W_x"]) + np.dot(h1m, w["W_h"]))
            zext = np.dot(np.tanh(np.dot(s, w["W_s"]) + hp), w["W_a"])
            beta = _softmax(np.concatenate((e, zext), axis=0), axis=0)[-1][0]
            chat = beta * s + (1 - beta) * ctx
            x2 = np.hstack((chat, h1.reshape(1, H)))
            h2, c2, z2, a2 = self._lstm(x2, h2m, c2m, w["lang_wi"], w["lang_wh"], w["lang_b"])
            # quirk B1: the explainer's logits use h2 only (explainers.py:1154)
            logits.append(np.dot(h2, w["output_w"]) + w["output_b"])
            for k, v in zip(keys1, [h1, c1, z1[0], z1[1], z1[2], z1[3], a1[0], a1[1], a1[2], a1[3]]):
                S[k].append(v.reshape(1, H))
            for k, v in zip(keys2, [h2, c2, z2[0], z2[1], z2[2], z2[3], a2[0], a2[1], a2[2], a2[3]]):
                S[k].append(v.reshape(1, H))
            S["ctx"].append(ctx)
            S["s"].append(s.reshape(1, H))
            S["chat"].append(chat.reshape(1, H))
            S["alpha"].append(alpha.reshape(1, self.L))
            S["beta"].append(np.reshape(beta, (1, 1)))
            x1s.append(x1)
            x2s.append(x2)
        self.S = {k: np.vstack(v) for k, v in S.items()}
        self.x1 = np.vstack(x1s)
        self.x2 = np.vstack(x2s)
        self.logits = np.vstack(logits)
        self.attention = self.S["alpha"]

    # ------------------------------------------------------------------ LRP
    def explain(self, t):
        """Relevance of logit(word t) (1-based) -> (R_F (1,s,s,D) float32, attention[t] (L,))."""
        if t > len(self.logits):
            raise NotImplementedError("index out of range of captions")
        return self._explain_adaptive(t) if self.kind == "adaptive" else self._explain_gridtd(t)

    def _features_back(self, r_glob, r_V):
        """Shared tail: global-feature dense, mean-pool split, image_features dense (explainers.py:634-659)."""
        w = self.w
        r_a = self._lin(r_glob, self.a, self.gp, w["global_w"])
        r_F = np.zeros((self.L, self.D), dtype="float32")
        if self.faithful:
            for l in range(self.L):
                r_F[l] = self._ew(r_a, self.F[l] / self.L, self.a)
                r_F[l] += self._lin(r_V[l], self.F[l], self.Vp[l], w["image_features_w"])
        else:
            r_F[:] = (self.F / self.L) * (r_a / stab(self.a))[None, :]
            r_F += self.F * np.dot(r_V / stab(self.Vp), w["image_features_w"].T)
        side = int(np.sqrt(self.L))
        return r_F.reshape(1, side, side, self.D)

    def _explain_adaptive(self, t):
        w, H, E, S = self.w, self.H, self.E, self.S
        k = self.caption[t - 1] - 1
        r_out = np.zeros((1, self.logits.shape[1]))
        r_out[0, k] = self.logits[t - 1, k]
        hc = S["h"][t] + S["chat"][t]
        if self.faithful:
            r_hc = self._lin(r_out, hc, self.logits[t - 1], w["output_w"])
        else:  # one non-zero output unit: the dense step collapses to a column of W_o
            r_hc = hc * w["output_w"][:, k] * (r_out[0, k] / stab(self.logits[t - 1, k:k + 1])[0])
        r_h = np.zeros((t + 1, H))
        r_c = np.zeros((t + 1, H))
        r_h[t] = self._ew(r_hc, S["h"][t], hc)
        r_chat = self._ew(r_hc, S["chat"][t], hc)
        beta = S["beta"][t][0]
        r_ctx = self._ew(r_chat, (1 - beta) * S["ctx"][t], S["chat"][t])
        r_c[t] = self._ew(r_chat, beta * S["s"][t], S["chat"][t])
        Wg = np.vstack((np.split(w["lstm_wi"], 4, 1)[2], np.split(w["lstm_wh"], 4, 1)[2]))
        xh = np.hstack((self.x[0:t], S["h"][0:t]))
        r_glob = np.zeros(E)
        r_word = np.zeros((t, E))
        for i in range(t)[::-1]:
            r_c[i + 1] += r_h[i + 1]
            r_g = self._ew(r_c[i + 1], S["ia"][i + 1] * np.tanh(S["zg"][i + 1]), S["c"][i + 1])
            r_c[i] = self._ew(r_c[i + 1], S["fa"][i + 1] * S["c"][i], S["c"][i + 1])
            r_xh = self._lin(r_g, xh[i], S["zg"][i + 1], Wg)
            r_h[i] = r_xh[2 * E:]
            r_glob += r_xh[E:2 * E]
            r_word[i] = r_xh[:E]
        r_V = np.zeros((self.L, H), dtype="float32")
        alpha = S["alpha"][t]
        if self.faithful:
            for l in range(self.L):
                r_V[l] = self._ew(r_ctx, self.Vf[l] * alpha[l], S["ctx"][t])
        else:
            r_V[:] = self.Vf * alpha[:, None] * (r_ctx / stab(S["ctx"][t]))[None, :]
        r_F = self._features_back(r_glob, r_V)
        rw = np.sum(r_word, axis=-1)
        rw[0] = 0
        m = np.max(np.abs(rw))
        if m:
            rw = rw / m
        self.r_words = rw[1:]
        return r_F, alpha

    def _explain_gridtd(self, t):
        w, H, E, S = self.w, self.H, self.E, self.S
        k = self.caption[t - 1] - 1
        r_out = np.zeros((1, self.logits.shape[1]))
        r_out[0, k] = self.logits[t - 1, k]
        hc = S["h2"][t] + S["chat"][t]
        if self.faithful:
            r_p = self._lin(r_out, hc, self.logits[t - 1], w["output_w"])
        else:
            r_p = hc * w["output_w"][:, k] * (r_out[0, k] / stab(self.logits[t - 1, k:k + 1])[0])
        r_c1, r_c2 = np.zeros((t + 1, H)), np.zeros((t + 1, H))
        r_h1, r_h2 = np.zeros((t + 1, H)), np.zeros((t + 1, H))
        r_chat = np.zeros((t, H))
        r_glob = np.zeros(E)
        r_word = np.zeros((t, E))
        r_V = np.zeros((self.L, H), dtype="float32")
        Wg1 = np.vstack((np.split(w["td_wi"], 4, 1)[2], np.split(w["td_wh"], 4, 1)[2]))
        Wg2 = np.vstack((np.split(w["lang_wi"], 4, 1)[2], np.split(w["lang_wh"], 4, 1)[2]))
        xh1 = np.hstack((self.x1[0:t], S["h1"][0:t]))
        xh2 = np.hstack((self.x2[0:t], S["h2"][0:t]))
        r_h2[t] = self._ew(r_p, S["h2"][t], hc)
        r_chat[t - 1] = self._ew(r_p, S["chat"][t], hc)
        for i in range(t)[::-1]:
            r_c2[i + 1] += r_h2[i + 1]
            r_g2 = self._ew(r_c2[i + 1], S["ia2"][i + 1] * np.tanh(S["zg2"][i + 1]), S["c2"][i + 1])
            r_c2[i] = self._ew(r_c2[i + 1], S["fa2"][i + 1] * S["c2"][i], S["c2"][i + 1])
            r_x2 = self._lin(r_g2, xh2[i], S["zg2"][i + 1], Wg2)
            r_h1[i + 1] += r_x2[H:2 * H]
            r_h2[i] += r_x2[2 * H:]
            r_chat[i] += r_x2[:H]
            beta = S["beta"][i + 1][0]
            r_s = self._ew(r_chat[i], beta * S["s"][i + 1], S["chat"][i + 1])
            r_ctx = self._ew(r_chat[i], S["ctx"][i + 1] * (1 - beta), S["chat"][i + 1])
            r_c1[i + 1] += r_s
            r_c1[i + 1] += r_h1[i + 1]
            r_g1 = self._ew(r_c1[i + 1], S["ia1"][i + 1] * np.tanh(S["zg1"][i + 1]), S["c1"][i + 1])
            r_c1[i] = self._ew(r_c1[i + 1], S["fa1"][i + 1] * S["c1"][i], S["c1"][i + 1])
            r_x1 = self._lin(r_g1, xh1[i], S["zg1"][i + 1], Wg1)
            r_h2[i] += r_x1[:H]
            r_glob += r_x1[H:H + E]
            r_word[i] = r_x1[H + E:H + 2 * E]
            alpha = S["alpha"][i + 1]
            if self.faithful:
                for l in range(self.L):
                    r_V[l] += self._ew(r_ctx, self.Vf[l] * alpha[l], S["ctx"][i + 1])
            else:
                r_V += self.Vf * alpha[:, None] * (r_ctx / stab(S["ctx"][i + 1]))[None, :]
            r_h1[i] += r_x1[H + 2 * E:]
        r_F = self._features_back(r_glob, r_V)
        self.r_words = np.sum(r_word, axis=-1)
        return r_F, S["alpha"][t]

    def explain_sentence(self):
        """explainers.py:183-189: words t = 1 .. len(caption)-1, attention[1:-1]."""
        rel = [self.explain(i + 1)[0] for i in range(len(self.caption) - 1)]
        return rel, self.attention[1:-1]

    # ------------------------------------------------------------------ gradient (frozen attention)
    def backward(self, t):
        return self._backward_adaptive(t) if self.kind == "adaptive" else self._backward_gridtd(t)

    def _cell_back(self, d_h, d_c_next, c_next, c_prev, acts):
        """One LSTM BPTT step as explainers.py:811-821 (float32 stores)."""
        ia, fa, ga, oa = acts
        f32 = np.float32
        d_oa = (d_h * np.tanh(c_next)).astype(f32)
        d_c = (d_c_next + d_h * oa * (1. - np.tanh(c_next) ** 2)).astype(f32)
        d_fa = (d_c * c_prev).astype(f32)
        d_c_prev = (d_c * fa).astype(f32)
        d_ia = (d_c * ga).astype(f32)
        d_ga = (d_c * ia).astype(f32)
        d_i = (d_ia * ia * (1 - ia)).astype(f32)
        d_f = (d_fa * fa * (1 - fa)).astype(f32)
        d_o = (d_oa * oa * (1 - oa)).astype(f32)
        d_g = (d_ga * (1 - ga ** 2)).astype(f32)
        return np.hstack((d_i[None], d_f[None], d_g[None], d_o[None])), d_c, d_c_prev

    def _backward_adaptive(self, t):
        w, H, E, S = self.w, self.H, self.E, self.S
        k = self.caption[t - 1] - 1
        d_hc = w["output_w"][:, k].astype(np.float64)
        d_h = np.zeros((t + 1, H), dtype="float32")
        d_c = np.zeros((t + 1, H), dtype="float32")
        d_h[t] = d_hc
        d_V = np.zeros((self.L, H), dtype="float32")
        d_V[:] = d_hc[None, :] * S["alpha"][t][:, None]     # no (1-beta), no sentinel path (quirk B4)
        d_V[self.Vf <= 0] = 0
        d_glob = np.zeros(H)
        d_words = np.zeros((t, E))
        for i in range(t)[::-1]:
            acts = (S["ia"][i + 1], S["fa"][i + 1], S["ga"][i + 1], S["oa"][i + 1])
            gates, d_c[i + 1], d_c[i] = self._cell_back(d_h[i + 1], d_c[i + 1], S["c"][i + 1], S["c"][i], acts)
            d_h[i] = np.dot(gates, w["lstm_wh"].T)
            d_x = np.dot(gates, w["lstm_wi"].T)[0].astype("float32")
            d_glob += d_x[E:]
            d_words[i] = d_x[:E]
        if self.g[0] <= 0:       # quirk B3: scalar test on element 0 masks all or nothing
            d_glob[:] = 0
        d_a = np.dot(d_glob, w["global_w"].T)
        d_F = np.zeros((self.L, self.D), dtype="float32")
        d_F[:] = (1.0 * d_a / self.L)[None, :]
        d_F += np.dot(d_V, w["image_features_w"].T)
        self.r_words = np.sum(d_words, axis=-1)
        side = int(np.sqrt(self.L))
        return d_F.reshape(1, side, side, self.D)

    def _backward_gridtd(self, t):
        w, H, E, S = self.w, self.H, self.E, self.S
        k = self.caption[t - 1] - 1
        d_p = w["output_w"][:, k].astype(np.float64)
        d_h1 = np.zeros((t + 1, H), dtype="float32")
        d_c1 = np.zeros((t + 1, H), dtype="float32")
        d_h2 = np.zeros((t + 1, H), dtype="float32")
        d_c2 = np.zeros((t + 1, H), dtype="float32")
        d_chat = np.zeros((t, H))
        d_V = np.zeros((self.L, H))
        d_glob = np.zeros((1, E))
        d_words = np.zeros((t, E))
        d_chat[t - 1] = d_p
        d_h2[t] = d_p
        for i in range(t)[::-1]:
            a2 = (S["ia2"][i + 1], S["fa2"][i + 1], S["ga2"][i + 1], S["oa2"][i + 1])
            g2, d_c2[i + 1], d_c2[i] = self._cell_back(d_h2[i + 1], d_c2[i + 1], S["c2"][i + 1], S["c2"][i], a2)
            d_h2[i] = np.dot(g2, w["lang_wh"].T)
            d_x2 = np.dot(g2, w["lang_wi"].T)[0]
            d_chat[i] += d_x2[:H]
            d_ctx = d_chat[i] * (1 - S["beta"][i + 1][0])
            d_h1[i + 1] += d_x2[H:]
            a1 = (S["ia1"][i + 1], S["fa1"][i + 1], S["ga1"][i + 1], S["oa1"][i + 1])
            g1, d_c1[i + 1], d_c1[i] = self._cell_back(d_h1[i + 1], d_c1[i + 1], S["c1"][i + 1], S["c1"][i], a1)
            d_h1[i] = np.dot(g1, w["td_wh"].T)
            d_x1 = np.dot(g1, w["td_wi"].T)[0]
            d_glob += d_x1[H:H + E]
            d_words[i] = d_x1[H + E:]
            d_V += d_ctx[None, :] * S["alpha"][i + 1][:, None]
            d_h2[i] += d_x1[:H]
        d_glob[0][self.g <= 0] = 0
        d_a = np.dot(d_glob, w["global_w"].T)
        d_V[self.Vf <= 0] = 0
        self.r_words = np.sum(d_words, axis=-1)
        d_F = np.zeros((self.L, self.D), dtype="float32")
        d_F[:] = np.dot(d_V, w["image_features_w"].T)
        d_F += (d_a[0] / self.L)[None, :]
        side = int(np.sqrt(self.L))
        return d_F.reshape(1, side, side, self.D)
