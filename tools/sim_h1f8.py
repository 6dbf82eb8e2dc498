"""CPU (numpy + torch dtypes) simulation of the operand arithmetic of the backward modes on one long contraction with a
heavy-tailed, mixed-sign message: relative RMS error of
  bf16x3   (a_hi + a_lo)(w_hi + w_lo) without lo*lo, bf16 parts                               3 product-equivalents
  f16x2    fp16(a) (w_hi + w_lo), fp16 parts                                                 2
  h1f8     fp16(a) w_hi  +  e4m3(fp16(a)) e4m3(w_lo)  +  e4m3((a - fp16(a)) 2^13) e4m3(w 2^-13)   1 + 2 x 0.5 = 2
This is the design argument for the fp16 + fp8 mode; the measured chain errors are in
profiles/r02_conv_modes.jsonl.  Usage: python tools/sim_h1f8.py"""
import numpy as np
import torch


def bf16(x): return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
def f16(x): return torch.from_numpy(x).to(torch.float16).to(torch.float32).numpy()
def e4m3(x): return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.float8_e4m3fn).to(torch.float32).numpy()


rng = np.random.default_rng(0)
for K, tail in ((4608, 1.5), (576, 1.5), (4608, 0.0), (2304, 2.5)):
    M, N = 256, 64
    A = (rng.standard_normal((M, K)) * np.exp(tail * rng.standard_normal((M, K)))).astype(np.float32)
    W = (rng.standard_normal((K, N)) * np.sqrt(2 / K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64)
    err = lambda y: float(np.sqrt(((y - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))   # noqa: E731
    ah = bf16(A); al = bf16(A - ah); wh = bf16(W); wl = bf16(W - wh)
    y3 = ah.astype(np.float64) @ wh + ah.astype(np.float64) @ wl + al.astype(np.float64) @ wh
    sc = 2.0 ** np.round(np.log2(16 / np.abs(A).max(axis=1, keepdims=True)))     # per-row power-of-two scale, maximum near 2^4
    As = (A * sc).astype(np.float32); a16 = f16(As)
    wsc = 2.0 ** 13 / (2 ** np.ceil(np.log2(np.abs(W).max())))                  # weights scaled into [2^12, 2^13)
    Ws = (W * wsc).astype(np.float32); wh16 = f16(Ws); wl16 = f16(Ws - wh16)
    y2 = (a16.astype(np.float64) @ (wh16.astype(np.float64) + wl16)) / sc / wsc
    r8 = e4m3((As - a16) * 2.0 ** 13); w8 = e4m3(Ws / 2.0 ** 13); a8 = e4m3(a16); wl8 = e4m3(wl16)
    y8 = (a16.astype(np.float64) @ wh16.astype(np.float64) + a8.astype(np.float64) @ wl8.astype(np.float64)
          + r8.astype(np.float64) @ w8.astype(np.float64)) / sc / wsc
    print("K=%d tail=%.1f   bf16x3 %.2e | f16x2 %.2e | h1f8 %.2e" % (K, tail, err(y3), err(y2), err(y8)))
