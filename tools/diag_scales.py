"""GPU: two-product backward -- where the stored fp16 message planes sit (max per word and layer vs the 2^target
prediction) and what the scale target does to the conservation sum.  Writes gpurun_out/diag_scales.jsonl."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lrp_imagecaptioning_b200 import synth                       # noqa: E402
from lrp_imagecaptioning_b200.encoder import ImageModel          # noqa: E402
from lrp_imagecaptioning_b200.analyzers import create_analyzer   # noqa: E402
from oracle import encoder_ref as ER                             # noqa: E402
from tests.util import linf_rel, l2_rel                          # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "diag_scales.jsonl")
hw, n = 224, 3
W = synth.vgg16_weights(0, bias_std=0.01)
x = synth.images(n, hw, 1)
idx = np.arange(n, dtype=np.int32)
rules = {"presetA": ("lrp.sequential_preset_a", {}, dict(epsilon=0.01)), "a2b1": ("lrp.alpha_2_beta_1", {}, {}),
         "eps": ("lrp.epsilon", dict(epsilon=0.01), dict(epsilon=0.01))}
ref, R = {}, None
for target in (4, 8, 11, 13):
    os.environ["LRPCAP_MSG_TARGET_EXP"] = str(target)
    for rule, (om, okw, akw) in rules.items():
        m = ImageModel(W, image_hw=hw, precision="f16x2")
        F = m.predict(x)
        if R is None:
            R = (F[idx] * np.random.default_rng(5).standard_normal((n,) + F.shape[1:])).astype(np.float32)
        got = create_analyzer(om, m, **akw).analyze_batch(x, idx, R).cpu().numpy()
        if rule not in ref:
            force = ER.Forced(m.pool_routes())
            ref[rule] = ER.analyze(om, x, R, W, force=force, **okw)
        mx, kt = m.message_scales()
        row = {"target_exp": target, "rule": rule, "finite": bool(np.isfinite(got).all()),
               "linf": [linf_rel(got[i], ref[rule][i]) for i in range(n)], "l2": [l2_rel(got[i], ref[rule][i]) for i in range(n)],
               "signed_sum_rel": [float((got[i].astype(np.float64).sum() - ref[rule][i].astype(np.float64).sum()) /
                                        np.abs(ref[rule][i]).astype(np.float64).sum()) for i in range(n)],
               "log2_stored_max_by_layer": [[round(float(np.log2(v)), 1) if v > 0 else None for v in mx[l]] for l in range(mx.shape[0])],
               "kt_by_layer": kt.tolist()}
        print(json.dumps(row), flush=True)
        with open(OUT, "a") as f:
            f.write(json.dumps(row) + "\n")
        m.close()
