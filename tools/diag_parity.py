"""GPU diagnostics behind the encoder parity gates:

  1. flips   -- per image: how many max-pool arg-max routes (and, for the mask rules, ReLU decisions) of the CUDA forward
                differ from the torch oracle's, and the map error against the oracle as-is and against the oracle pinned
                to the CUDA forward's decisions (oracle/encoder_ref.py: Forced).
  2. sums    -- relevance-conservation drift of the alpha-beta family as a function of the accumulator promotion interval
                of the transposed-conv GEMMs (tensor-core fp32 accumulation truncates; same-sign chains drift).
  3. trunc   -- the truncation itself on one long all-positive contraction.

Writes gpurun_out/diag_parity.jsonl.  Usage: python tools/diag_parity.py [flips] [sums] [trunc] [--hw 224] [--images 3]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lrp_imagecaptioning_b200 import synth, _lib                      # noqa: E402
from lrp_imagecaptioning_b200.encoder import ImageModel               # noqa: E402
from lrp_imagecaptioning_b200.analyzers import create_analyzer        # noqa: E402
from oracle import encoder_ref as ER                                  # noqa: E402
from tests.util import linf_rel, l2_rel, sum_err, topk_cells          # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "diag_parity.jsonl")
RULES = {
    "eps": ("lrp.epsilon", dict(epsilon=0.01), "lrp.epsilon", dict(epsilon=0.01), False),
    "presetA": ("lrp.sequential_preset_a", {}, "lrp.sequential_preset_a", dict(epsilon=0.01), False),
    "a2b1": ("lrp.alpha_2_beta_1", {}, "lrp.alpha_2_beta_1", {}, False),
    "zplus": ("lrp.z_plus", {}, "lrp.z_plus", {}, False),
    "z": ("lrp.z", {}, "lrp.z", {}, True),
    "gradient": ("gradient", {}, "gradient", {}, True),
    "guided": ("guided_backprop", {}, "guided_backprop", {}, True),
    "ixg": ("input_t_gradient", {}, "input_t_gradient", {}, True),
}


def emit(row):
    print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "a") as f:
        f.write(json.dumps(row) + "\n")


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def forced_from(model, masks):
    routes = model.pool_routes()
    mk = None
    if masks:
        mk = {l: model.multiplier(l) != 0 for l in range(12)}
    return ER.Forced(routes, mk), routes


def flips(hw, n, rules, precision="bf16x3"):
    W = synth.vgg16_weights(0, bias_std=0.01)
    x = synth.images(n, hw, 1)
    idx = np.arange(n, dtype=np.int32)
    own_routes = ER.pool_routes(x, W)
    for rule in rules:
        om, okw, an, akw, masks = RULES[rule]
        m = ImageModel(W, image_hw=hw, precision=precision)
        F = m.predict(x)
        R = (F[idx] * np.random.default_rng(5).standard_normal((n,) + F.shape[1:])).astype(np.float32)
        a = create_analyzer(an, m, **akw)
        got = a.analyze_batch(x, idx, R).cpu().numpy()
        force, routes = forced_from(m, masks)
        t0 = time.time()
        ref = ER.analyze(om, x, R, W, **okw)
        ref_f = ER.analyze(om, x, R, W, force=force, **okw)
        dt = time.time() - t0
        for i in range(n):
            nflip = {int(l): int((routes[l][i] != own_routes[l][i]).sum()) for l in routes}
            emit({"what": "flips", "rule": rule, "hw": hw, "precision": precision, "image": i, "route_flips": nflip,
                  "linf_vs_oracle": linf_rel(got[i], ref[i]), "l2_vs_oracle": l2_rel(got[i], ref[i]),
                  "linf_vs_pinned": linf_rel(got[i], ref_f[i]), "l2_vs_pinned": l2_rel(got[i], ref_f[i]),
                  "sum_vs_oracle": sum_err(got[i], ref[i]), "sum_vs_pinned": sum_err(got[i], ref_f[i]),
                  "top10_same_pinned": topk_cells(got[i], 10) == topk_cells(ref_f[i], 10),
                  "top10_same_oracle": topk_cells(got[i], 10) == topk_cells(ref[i], 10), "oracle_s": dt})
        m.close()


def sums(hw, n, rules, promotes):
    W = synth.vgg16_weights(0, bias_std=0.01)
    x = synth.images(n, hw, 1)
    idx = np.arange(n, dtype=np.int32)
    for rule in rules:
        om, okw, an, akw, masks = RULES[rule]
        ref = None
        for prec, p in [("fp32", 0)] + [("bf16x3", p) for p in promotes]:
            m = ImageModel(W, image_hw=hw, precision=prec)
            F = m.predict(x)
            if ref is None:
                R = (F[idx] * np.random.default_rng(5).standard_normal((n,) + F.shape[1:])).astype(np.float32)
                ref = ER.analyze(om, x, R, W, **okw)
            if prec != "fp32":
                m.set_promote(p)
            got = create_analyzer(an, m, **akw).analyze_batch(x, idx, R).cpu().numpy()
            emit({"what": "sums", "rule": rule, "hw": hw, "precision": prec, "promote": p,
                  "sum_err": [sum_err(got[i], ref[i]) for i in range(n)],
                  "signed_rel": [float((got[i].astype(np.float64).sum() - ref[i].astype(np.float64).sum()) /
                                       np.abs(ref[i]).astype(np.float64).sum()) for i in range(n)],
                  "l2": [l2_rel(got[i], ref[i]) for i in range(n)]})
            m.close()


def trunc():
    g = np.random.default_rng(0)
    for C in (64, 256, 512):
        A = g.uniform(0.5, 1.5, size=(1, 16, 16, C)).astype(np.float32)
        B = g.uniform(0.5, 1.5, size=(9, C, 64)).astype(np.float32)
        import torch
        ref = torch.nn.functional.conv2d(torch.from_numpy(A).double().permute(0, 3, 1, 2),
                                         torch.from_numpy(B.reshape(3, 3, C, 64)).double().permute(3, 2, 0, 1), padding=1)
        ref = ref.permute(0, 2, 3, 1).numpy()
        row = {"what": "trunc", "C": C, "K": 9 * C}
        for name, prec in (("fp32_simt", 0), ("bf16x2_nopromote", 1), ("bf16x3planes_promoted", 2), ("f16x2_promoted", 3),
                           ("h1x2_two_product_nopromote", 4)):
            out = _lib.debug_conv(prec, A, B, 9).astype(np.float64)
            if prec == 4:    # the A operand is one fp16 plane by construction: compare against the contraction of that operand
                A16 = A.astype(np.float16).astype(np.float32)
                r16 = torch.nn.functional.conv2d(torch.from_numpy(A16).double().permute(0, 3, 1, 2),
                                                 torch.from_numpy(B.reshape(3, 3, C, 64)).double().permute(3, 2, 0, 1), padding=1)
                rel16 = (out - r16.permute(0, 2, 3, 1).numpy()) / ref
                row[name + "_vs_fp16_operand"] = {"mean_rel": float(rel16[:, 4:12, 4:12].mean())}
            rel = (out - ref) / ref
            row[name] = {"mean_rel": float(rel[:, 4:12, 4:12].mean()), "rms_rel": float(np.sqrt((rel[:, 4:12, 4:12] ** 2).mean()))}
        emit(row)


if __name__ == "__main__":
    hw = arg("--hw", 224)
    n = arg("--images", 3)
    what = [a for a in sys.argv[1:] if a in ("flips", "sums", "trunc")] or ["trunc", "sums", "flips"]
    rules = arg("--rules", "")
    if "trunc" in what:
        trunc()
    if "sums" in what:
        sums(hw, n, (rules.split(",") if rules else ["zplus", "a2b1", "presetA"]), [0, 8, 2])
    if "flips" in what:
        flips(hw, n, (rules.split(",") if rules else ["eps", "presetA", "a2b1", "z", "gradient", "guided"]),
              precision=arg("--precision", "bf16x3"))
