"""GPU: why small images sit further from the oracle in the two-product modes.  Per image: active units per layer,
error of every precision against the pinned oracle, and how uniform the error is (median got/ref on the large cells)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lrp_imagecaptioning_b200 import synth
from lrp_imagecaptioning_b200.encoder import ImageModel
from lrp_imagecaptioning_b200.analyzers import create_analyzer
from oracle import encoder_ref as ER
from tests.util import linf_rel, l2_rel

W = synth.vgg16_weights(0, bias_std=0.01)
n = 5
idx = np.arange(n, dtype=np.int32)
for hw in (32, 64):
    x = synth.images(n, hw, 1)
    ref = None
    for prec in ("bf16x3", "f16x2", "h1f8"):
        m = ImageModel(W, image_hw=hw, precision=prec)
        F = m.predict(x)
        if ref is None:
            R = (F[idx] * np.random.default_rng(2).standard_normal((n,) + F.shape[1:])).astype(np.float32)
            force = ER.Forced(m.pool_routes())
            ref = ER.analyze("lrp.alpha_1_beta_0", x, R, W, force=force)
            act = [[int((m.multiplier(l)[i] != 0).sum()) for l in range(12)] for i in range(n)]
            print(json.dumps({"hw": hw, "nonzero_features": [int((F[i] != 0).sum()) for i in range(n)], "active_multipliers_by_layer": act}))
        got = create_analyzer("lrp.alpha_1_beta_0", m).analyze_batch(x, idx, R).cpu().numpy()
        row = {"hw": hw, "precision": prec, "linf": [], "l2": [], "median_ratio_minus_1": []}
        for i in range(n):
            big = np.abs(ref[i]) > 0.1 * np.abs(ref[i]).max()
            row["linf"].append(linf_rel(got[i], ref[i])); row["l2"].append(l2_rel(got[i], ref[i]))
            row["median_ratio_minus_1"].append(float(np.median(got[i][big] / ref[i][big]) - 1))
        print(json.dumps(row), flush=True)
        m.close()
