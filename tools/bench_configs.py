"""Secondary measurements for the parity-test configurations of BASELINE.json (not the bench line): explained words/s on one
GPU for  config 3 (adaptive attention, V = 10 000, alpha-beta encoder rules)  and  config 4 (grid-TD method sweep:
LRP-eps, Gradient x Input, Guided backprop, Grad-CAM, Guided-Grad-CAM).  Writes gpurun_out/bench_configs.json.
Usage (GPU box): python tools/bench_configs.py [--images 64] [--steps 3]"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import _lib, synth, gradcam
from lrp_imagecaptioning_b200.encoder import RuleSpec
from lrp_imagecaptioning_b200.engine import ExplainEngine, word_list, METHOD_LRP, METHOD_GRADIENT
from lrp_imagecaptioning_b200.model import CaptioningModel

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
T, HW, V = 20, 224, 10000
x = torch.from_numpy(synth.images(args.images, HW, 100)).cuda()
wi, wt = word_list(args.images, T)


def timed(fn):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    return {"ms_per_step": ms, "words_per_s": len(wi) / (ms / 1e3)}


out = {"images": args.images, "words_per_image": T, "vocab": V, "results": {}}
for kind, cases in (("adaptive", [("config3 adaptive, LRP decoder + PresetA (alpha1 beta0, bias)", METHOD_LRP, RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True), None),
                                  ("config3 adaptive, LRP decoder + alpha2 beta1", METHOD_LRP, RuleSpec(_lib.RULE_ALPHA_BETA, alpha=2, beta=1, bias=True), None)]),
                    ("gridtd", [("config4 grid-TD, LRP decoder + LRPEpsilon(0.01)", METHOD_LRP, RuleSpec(_lib.RULE_EPSILON, epsilon=0.01, bias=True), None),
                                ("config4 grid-TD, gradient decoder + Gradient x Input", METHOD_GRADIENT, RuleSpec(_lib.RULE_INPUT_T_GRADIENT), None),
                                ("config4 grid-TD, gradient decoder + GuidedBackprop", METHOD_GRADIENT, RuleSpec(_lib.RULE_GUIDED_BACKPROP), None),
                                ("config4 grid-TD, Grad-CAM (decoder gradient only)", METHOD_GRADIENT, RuleSpec(_lib.RULE_GRADIENT), "cam"),
                                ("config4 grid-TD, Guided-Grad-CAM", METHOD_GRADIENT, RuleSpec(_lib.RULE_GUIDED_BACKPROP), "guided_cam")])):
    model = CaptioningModel.synthetic(kind, vocab_size=V, image_hw=HW, seed=0)
    model.image_model.set_chunk_words(320)
    for name, method, rule, cam in cases:
        eng = ExplainEngine(model, rule=rule)

        def step():
            eng.forward(x, T=T, greedy=True)
            if cam is None:
                return eng.explain_words(wi, wt, method=method)
            R_head, _ = eng.decoder.backward(wi, wt, want_words=False)
            feats = model.image_model.features()
            c = gradcam.grad_cam_batch(feats.reshape(feats.shape[0], -1, feats.shape[-1]), wi, R_head)
            if cam == "cam":
                return c
            maps = model.image_model.relevance(wi, R_head.view(-1, HW // 16, HW // 16, R_head.shape[-1]))
            return gradcam.scale_maps(maps, c)
        t0 = time.time()
        out["results"][name] = timed(step)
        print(name, out["results"][name], "(%.1f s)" % (time.time() - t0), flush=True)
    model.image_model.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)
