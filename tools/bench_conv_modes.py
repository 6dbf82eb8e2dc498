"""GPU micro-benchmark of the backward arithmetic modes on the encoder relevance alone (no decoder): 4 images x 80 words
(one 320-word chunk) at 224 x 224, epsilon rule.  Prints ms per chunk and the per-image error against the pinned oracle."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import synth, _lib
from lrp_imagecaptioning_b200.encoder import ImageModel, RuleSpec
from oracle import encoder_ref as ER
from tests.util import linf_rel, l2_rel, sum_err

rules = {"eps": (RuleSpec(_lib.RULE_EPSILON, epsilon=0.01), "lrp.epsilon", dict(epsilon=0.01)),
         "presetA": (RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True), "lrp.sequential_preset_a", {}),
         "gradient": (RuleSpec(_lib.RULE_GRADIENT), "gradient", {})}
W = synth.vgg16_weights(0, bias_std=0.01)
n_img, per = 4, 80
x = synth.images(n_img, 224, 1)
idx = np.repeat(np.arange(n_img), per).astype(np.int32)
for rname in (sys.argv[1:] or ["eps", "presetA"]):
    rule, om, okw = rules[rname]
    ref = None
    for prec in ("bf16x3", "f16x2", "h1f8"):
        m = ImageModel(W, image_hw=224, precision=prec)
        m.set_chunk_words(n_img * per)
        m.forward(x, rule)
        F = m.features()
        g = torch.Generator(device="cuda").manual_seed(0)
        R = (F[torch.as_tensor(idx, device=F.device).long()] * torch.randn((len(idx),) + tuple(F.shape[1:]), device=F.device, generator=g)).contiguous()
        out = m.relevance(idx, R)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            out = m.relevance(idx, R)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        if ref is None:   # pinned oracle for word 0 of each image
            force = ER.Forced(m.pool_routes(), {l: m.multiplier(l) != 0 for l in range(12)} if rname == "gradient" else None)
            sel = np.arange(n_img) * per
            ref = ER.analyze(om, x, R[torch.as_tensor(sel, device=R.device)].cpu().numpy(), W, force=force, **okw)
        got = out[torch.as_tensor(np.arange(n_img) * per, device=out.device)].cpu().numpy()
        row = {"rule": rname, "precision": prec, "ms_per_320_words": ms, "finite": bool(np.isfinite(got).all()),
               "linf": [linf_rel(got[i], ref[i]) if np.isfinite(got[i]).all() else None for i in range(n_img)],
               "l2": [l2_rel(got[i], ref[i]) if np.isfinite(got[i]).all() else None for i in range(n_img)],
               "sum": [sum_err(got[i], ref[i]) if np.isfinite(got[i]).all() else None for i in range(n_img)]}
        print(json.dumps(row), flush=True)
        m.close()
