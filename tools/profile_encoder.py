"""Short encoder-only workload for ncu: forward of 4 images, then the relevance backward for 4 x WORDS_PER_IMAGE words
(default 80 each = one 320-word chunk, the bench's chunk) at 224x224.  N_IMAGES=64 WORDS_PER_IMAGE=1 profiles the forward
of a bench-sized batch.
  RULE=eps|presetA|a2b1 (default eps)   PRECISION=tc|bf16x3|f16x2 (default tc)
`ncu -k regex:"tc_conv|last_dgrad" -s <forward launches> -c 13` captures the 12 transposed-conv launches + the last conv."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import synth, _lib
from lrp_imagecaptioning_b200.encoder import ImageModel, RuleSpec

n_img, per = int(os.environ.get("N_IMAGES", "4")), int(os.environ.get("WORDS_PER_IMAGE", "80"))
rule = {"eps": RuleSpec(_lib.RULE_EPSILON, epsilon=0.01), "presetA": RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True),
        "a2b1": RuleSpec(_lib.RULE_ALPHA_BETA, alpha=2, beta=1, bias=True)}[os.environ.get("RULE", "eps")]
m = ImageModel(synth.vgg16_weights(0), image_hw=224, precision=os.environ.get("PRECISION", "tc"))
m.set_chunk_words(n_img * per)
x = synth.images(n_img, 224, 1)
m.forward(x, rule)
F = m.features()
idx = np.repeat(np.arange(n_img), per).astype(np.int32)
R = (F[torch.as_tensor(idx, device=F.device).long()] * torch.randn((len(idx),) + tuple(F.shape[1:]), device=F.device)).contiguous()
out = m.relevance(idx, R)
torch.cuda.synchronize()
print("ok", float(out.abs().max()), "launches", m.launches())
