"""Short decoder-only workload for ncu: grid-TD (or KIND=adaptive) decoder, V = 10 000, 64 feature grids, greedy 20-step
forward, then the relevance of all 1 280 words.  The element-wise rule kernels of BASELINE.json north_star (b):
  gate pass-through            lrp_cell_kernel                 (explainers.py:604-619)
  sentinel / context split     lrp_scatter_lang_kernel, lrp_init_kernel   (:583-602, :1252-1269)
  attention redistribution     uv_gridtd_rows_f32_kernel / uv_adaptive_kernel   (:648-659, :1292-1299)
  grid-feature assembly        final_kernel                    (:641-659)
  forward attention / context  fwd_scores_kernel, fwd_ctx_kernel   (:413-421)
`ncu --set full -k regex:"lrp_cell|lrp_scatter|uv_|final_kernel|fwd_ctx|fwd_scores|fwd_lstm" -c 40`."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import synth
from lrp_imagecaptioning_b200.decoder import DecoderEngine

kind = os.environ.get("KIND", "gridtd")
N, T, V = 64, 20, 10000
dec = synth.decoder_weights(kind, V=V, seed=1)
F = torch.from_numpy(synth.features(N, seed=2)).cuda()
eng = DecoderEngine(dec)
cap = eng.forward(F, T=T, greedy=True, eos=2)
wi = np.repeat(np.arange(N), T).astype(np.int32)
wt = np.tile(np.arange(1, T + 1), N).astype(np.int32)
R, _, _ = eng.relevance(wi, wt, want_words=False, want_attention=False)
torch.cuda.synchronize()
print("ok", float(R.abs().max()), "launches", eng.launches())
