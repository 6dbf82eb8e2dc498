"""GPU: where the decoder-LRP error of the full-size synthetic case comes from -- the same case under the forward variants
(fused tensor-core forward / per-GEMM tensor-core forward / all-fp64 forward)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    from oracle.decoder_ref import DecoderRef
    from tests.util import linf_rel, l2_rel
    kind = sys.argv[2]
    cfg = dict(V=10000, H=512, E=512, D=512, L=196, T=20, N=2)
    dec = synth.decoder_weights(kind, V=cfg["V"], H=cfg["H"], E=cfg["E"], D=cfg["D"], seed=11)
    F = synth.features(cfg["N"], L=cfg["L"], D=cfg["D"], seed=12)
    cap = synth.captions(cfg["N"], cfg["T"], cfg["V"], seed=13)
    eng = DecoderEngine(dec)
    eng.forward(F, cap)
    N, T = cfg["N"], cfg["T"]
    wi = np.repeat(np.arange(N), T).astype(np.int32)
    wt = np.tile(np.arange(1, T + 1), N).astype(np.int32)
    R = eng.relevance(wi, wt)[0].cpu().numpy()
    G = eng.backward(wi, wt)[0].cpu().numpy()
    errs, gerrs = [], []
    for n in range(N):
        o = DecoderRef(dec).forward(F[n], list(cap[n]))
        for t in range(1, T + 1):
            rF, _ = o.explain(t)
            errs.append(linf_rel(R[n * T + t - 1], rF.reshape(196, 512)))
            gerrs.append(linf_rel(G[n * T + t - 1], o.backward(t).reshape(196, 512)))
    print(json.dumps({"kind": kind, "env": {k: v for k, v in os.environ.items() if k.startswith("LRPCAP_")},
                      "lrp_linf_max": max(errs), "lrp_linf_sorted_top5": sorted(errs)[-5:], "grad_linf_max": max(gerrs),
                      "grad_linf_sorted_top5": sorted(gerrs)[-5:]}))
else:
    for env in ({}, {"LRPCAP_DECODER_FUSED": "0"}, {"LRPCAP_DECODER_TC_FWD": "0"}, {"LRPCAP_DECODER_FP64_GEMM": "1", "LRPCAP_DECODER_TC_FWD": "0"}):
        for kind in ("gridtd", "adaptive"):
            e = dict(os.environ); e.update(env)
            r = subprocess.run([sys.executable, __file__, "child", kind], env=e, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-400:], flush=True)
