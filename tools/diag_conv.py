"""GPU diagnostic for the tcgen05 conv kernel: runs lrpcap_debug_conv over a battery of shapes and structured
probes, compares with a float64 numpy reference, writes gpurun_out/diag_conv.json (+ small dumps)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import _lib

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def ref_conv(A, B, taps):
    A = A.astype(np.float64); B = B.astype(np.float64)
    items, H, W, C = A.shape
    out = np.zeros((items, H, W, B.shape[-1]))
    if taps == 1:
        return A @ B[0]
    Ap = np.pad(A, ((0, 0), (1, 1), (1, 1), (0, 0)))
    for t in range(9):
        dy, dx = t // 3, t % 3
        out += Ap[:, dy:dy + H, dx:dx + W, :] @ B[t]
    return out


def stats(got, ref):
    bad = ~np.isfinite(got)
    err = np.abs(np.where(bad, 0, got) - ref)
    scale = np.max(np.abs(ref)) + 1e-30
    return {"max_rel_to_max": float(err.max() / scale), "nan": int(bad.sum()), "n": int(got.size),
            "mean_rel": float(err.mean() / scale)}


report = {"cases": []}
rng = np.random.default_rng(0)
cases = [  # (items, H, W, C, Nout, taps)
    (1, 8, 16, 64, 64, 1), (1, 8, 16, 64, 64, 9), (2, 16, 16, 64, 64, 9), (1, 16, 32, 128, 128, 9),
    (1, 14, 14, 128, 256, 9), (1, 14, 14, 512, 512, 9), (3, 28, 28, 256, 256, 9), (1, 56, 56, 128, 64, 9),
    (1, 4, 4, 64, 128, 9), (1, 2, 2, 512, 512, 9), (2, 32, 32, 64, 64, 9), (5, 7, 7, 64, 64, 1), (1, 224, 224, 64, 64, 9),
    (1, 32, 48, 128, 64, 9), (3, 48, 32, 256, 128, 9), (150, 16, 16, 64, 64, 9),
]
for prec, pname in ((_lib.PREC_FP32_SIMT, "simt"), (_lib.PREC_BF16X3_TC, "tc"), (2, "tc3")):
    for (items, H, W, C, Nout, taps) in cases:
        A = rng.standard_normal((items, H, W, C)).astype(np.float32)
        B = (rng.standard_normal((taps, C, Nout)) / np.sqrt(taps * C)).astype(np.float32)
        t0 = time.time()
        try:
            got = _lib.debug_conv(prec, A, B, taps)
            st = stats(got, ref_conv(A, B, taps))
        except Exception as e:  # noqa
            st = {"error": repr(e)}
        st.update({"impl": pname, "shape": [items, H, W, C, Nout, taps], "sec": round(time.time() - t0, 3)})
        print(st, flush=True)
        report["cases"].append(st)

# structured probe: 1x1 conv with identity weights -> out must equal A (reveals swizzle / layout permutations)
H, W, C = 8, 16, 64
A = (np.arange(H * W)[:, None] * 1.0 + np.arange(C)[None, :] / 128.0).reshape(1, H, W, C).astype(np.float32)
B = np.eye(C, dtype=np.float32)[None]
try:
    got = _lib.debug_conv(_lib.PREC_BF16X3_TC, A, B, 1)
    report["identity_probe"] = stats(got, A.astype(np.float64))
    np.save(os.path.join(OUT, "identity_probe_out.npy"), got[0].reshape(H * W, C)[:32])
    # tap probe: 3x3 conv, only tap t non-zero = identity -> out = shifted A
    taps_ok = []
    for t in range(9):
        B9 = np.zeros((9, C, C), dtype=np.float32); B9[t] = np.eye(C)
        g = _lib.debug_conv(_lib.PREC_BF16X3_TC, A, B9, 9)
        taps_ok.append(stats(g, ref_conv(A, B9, 9))["max_rel_to_max"])
    report["tap_probe"] = taps_ok
    # precision probe: hi/lo split accuracy on a long K
    A2 = rng.standard_normal((1, 16, 16, 512)).astype(np.float32)
    B2 = (rng.standard_normal((9, 512, 64)) / 68.0).astype(np.float32)
    r2 = ref_conv(A2, B2, 9)
    report["precision_probe"] = {"tc": stats(_lib.debug_conv(_lib.PREC_BF16X3_TC, A2, B2, 9), r2),
                                 "tc3": stats(_lib.debug_conv(2, A2, B2, 9), r2),
                                 "simt": stats(_lib.debug_conv(_lib.PREC_FP32_SIMT, A2, B2, 9), r2)}
except Exception as e:  # noqa
    report["probe_error"] = repr(e)
print(json.dumps({k: v for k, v in report.items() if k != "cases"}, indent=1))
json.dump(report, open(os.path.join(OUT, "diag_conv.json"), "w"), indent=1)
