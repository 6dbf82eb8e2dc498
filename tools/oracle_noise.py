"""CPU study of how well-conditioned the encoder relevance rules are on the synthetic (random-init) VGG16.

(1) The fp32 torch oracle (oracle/encoder_ref.py) against a float64 emulation of the same rule chain: how far the
    *oracle itself* moves when its arithmetic changes -- the floor for any parity test against it.
(2) The rule chain in float64 with the GEMM operands of the forward / backward pass rounded to 16 significant bits
    (the bf16 hi+lo split): shows that the per-image forward needs fp32-exact operands (discrete ReLU / arg-max
    decisions, x/stab(z) quotients) while the per-word backward does not.

    python tools/oracle_noise.py [hw ...]      (default: 64 224; writes profiles/r01_oracle_noise.txt)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lrp_imagecaptioning_b200 import synth  # noqa: E402
from oracle import encoder_ref as ER  # noqa: E402

POOL = (1, 3, 6, 9)


def q16(t):
    hi = t.float().bfloat16().double()
    lo = (t.double() - hi).float().bfloat16().double()
    return hi + lo


def chain(rule, x, R, W, idx, eps=0.01, fwd_q=False, bwd_q=False):
    """float64 emulation of the library's multiplier chain; *_q: 16-bit GEMM operands."""
    dt = torch.float64
    X = torch.from_numpy(x).permute(0, 3, 1, 2).to(dt)
    G, cur, Ms = [], X, None
    for l, (k, b) in enumerate(W):
        w = torch.from_numpy(k).permute(3, 2, 0, 1).to(dt)
        bb = torch.from_numpy(b).to(dt)
        a_in = q16(cur) if (fwd_q and l > 0) else cur
        z = Fn.conv2d(a_in, q16(w) if (fwd_q and l > 0) else w, bb, padding=1)
        xo = torch.relu(z)
        if rule == "eps":
            d = z + torch.where(z >= 0, eps, -eps)
        elif rule == "grad":
            d = None
        else:
            wp = w * (w >= 0)
            if l == 0:
                wn = w * (w < 0)
                za = Fn.conv2d(cur * (cur >= 0), wp, None, padding=1) + Fn.conv2d(cur * (cur < 0), wn, None, padding=1) + bb[None, :, None, None]
            else:
                za = Fn.conv2d(a_in, q16(wp) if fwd_q else wp, bb, padding=1)
            d = za + (za == 0) * 1e-7
        g = (z > 0).to(dt) if d is None else xo / d
        m = g if d is None else 1 / d
        if l in POOL:
            p, ind = Fn.max_pool2d(xo, 2, 2, return_indices=True)
            mask = torch.zeros_like(xo).flatten(2)
            mask.scatter_(2, ind.flatten(2), 1.0)
            g = g * mask.view_as(xo)
            cur = p
        else:
            cur = xo
        G.append(g.float().double())
        Ms = m.float().double()
    s = torch.from_numpy(R).permute(0, 3, 1, 2).to(dt) * Ms[idx]
    for l in range(12, 0, -1):
        w = torch.from_numpy(W[l][0]).permute(3, 2, 0, 1).to(dt)
        if rule == "a1b0":
            w = w * (w >= 0)
        c = Fn.conv_transpose2d(q16(s), q16(w), padding=1) if bwd_q else Fn.conv_transpose2d(s, w, padding=1)
        if (l - 1) in POOL:
            c = c.repeat_interleave(2, 2).repeat_interleave(2, 3)
        s = c * G[l - 1][idx]
    w = torch.from_numpy(W[0][0]).permute(3, 2, 0, 1).to(dt)
    if rule == "a1b0":
        ca = Fn.conv_transpose2d(s, w * (w >= 0), padding=1)
        cb = Fn.conv_transpose2d(s, w * (w < 0), padding=1)
        xi = X[idx]
        out = torch.where(xi >= 0, xi * ca, xi * cb)
    else:
        c = Fn.conv_transpose2d(s, w, padding=1)
        out = c if rule == "grad" else X[idx] * c
    return out.permute(0, 2, 3, 1).numpy(), cur.permute(0, 2, 3, 1).numpy()


def linf(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


def l2(a, b):
    return float(np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.ravel()))


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 224]
    torch.set_num_threads(os.cpu_count() or 1)
    lines = []

    def say(s):
        print(s, flush=True)
        lines.append(s)
    W = synth.vgg16_weights(0, bias_std=0.01)
    for hw in sizes:
        x = synth.images(1, hw, 1)
        idx = np.array([0])
        F = ER.features(x, W)
        R = (F[idx] * np.random.default_rng(5).standard_normal(F[idx].shape)).astype(np.float32)
        for rule, om, kw in (("eps", "lrp.epsilon", dict(epsilon=0.01)), ("a1b0", "lrp.alpha_1_beta_0", {}), ("grad", "gradient", {})):
            truth, _ = chain(rule, x, R, W, idx)
            ref = ER.analyze(om, x[idx], R, W, **kw)
            say("hw=%3d %-5s fp32 oracle vs float64 chain       : linf %.2e  l2 %.2e" % (hw, rule, linf(ref, truth), l2(ref, truth)))
            for fq, bq, name in ((True, False, "forward 16-bit operands "), (False, True, "backward 16-bit operands")):
                o, _ = chain(rule, x, R, W, idx, fwd_q=fq, bwd_q=bq)
                say("hw=%3d %-5s %s vs float64 chain: linf %.2e  l2 %.2e" % (hw, rule, name, linf(o, truth), l2(o, truth)))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r01_oracle_noise.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
