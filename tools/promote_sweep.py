"""GPU: accuracy of the encoder relevance at 224x224 vs the torch oracle as a function of the backward promotion
interval (tensor-core accumulator -> fp32 registers every n k-steps)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import synth
from lrp_imagecaptioning_b200.encoder import ImageModel
from lrp_imagecaptioning_b200.analyzers import create_analyzer
from oracle import encoder_ref as ER

W = synth.vgg16_weights(0, bias_std=0.01)
x = synth.images(1, 224, 1)
idx = np.array([0, 0], dtype=np.int32)
m32 = ImageModel(W, image_hw=224, precision="fp32")
F = m32.predict(x)
R = (F[idx] * np.random.default_rng(5).standard_normal((2,) + F.shape[1:])).astype(np.float32)
def err(a, b): return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max()), float(np.linalg.norm((a.astype(np.float64) - b).ravel()) / np.linalg.norm(b.ravel()))
out = []
for rule, om, kw, an, akw in (("eps", "lrp.epsilon", dict(epsilon=0.01), "lrp.epsilon", dict(epsilon=0.01)),
                              ("presetA", "lrp.sequential_preset_a", {}, "lrp.sequential_preset_a", dict(epsilon=0.01)),
                              ("guided", "guided_backprop", {}, "guided_backprop", {})):
    ref = ER.analyze(om, x[idx], R, W, **kw).astype(np.float64)
    f32 = create_analyzer(an, m32, **akw).analyze_batch(x, idx, R).cpu().numpy()
    row = {"rule": rule, "fp32_vs_oracle": err(f32, ref)}
    for p in (0, 1, 2, 4, 8):
        m = ImageModel(W, image_hw=224, precision="bf16x3")
        m.set_promote(p)
        got = create_analyzer(an, m, **akw).analyze_batch(x, idx, R).cpu().numpy()
        row["promote_%d_vs_oracle" % p] = err(got, ref)
        row["promote_%d_vs_fp32mode" % p] = err(got, f32.astype(np.float64))
        m.close()
    print(json.dumps(row), flush=True)
    out.append(row)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "promote_sweep.json"), "w"), indent=1)
