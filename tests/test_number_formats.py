"""CPU model of the backward operand formats (csrc/epilogue.cuh: StoreSplit / StoreH1 / StoreH1F8,
csrc/encoder_kernels.cu: prep_weights): what each one keeps of an fp32 message x weight contraction.  Not a kernel test --
it pins the number-format argument (three bf16 products ~ 4e-6, fp16 + fp8 ~ 1e-5, one fp16 message plane ~ 2e-4) that the
precision policy of `PREC_TC_AUTO` rests on, on the CPU, with the same splits the CUDA code stores."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _bf16(x): return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
def _f16(x): return torch.from_numpy(x).to(torch.float16).to(torch.float32).numpy()
def _e4m3(x): return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.float8_e4m3fn).to(torch.float32).numpy()


def _errors(K, tail, seed):
    rng = np.random.default_rng(seed)
    M, N = 128, 64
    A = (rng.standard_normal((M, K)) * np.exp(tail * rng.standard_normal((M, K)))).astype(np.float32)   # heavy-tailed, mixed sign
    W = (rng.standard_normal((K, N)) * np.sqrt(2 / K)).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64)
    err = lambda y: float(np.sqrt(((y - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))   # noqa: E731
    # three products on bf16 planes: (a_hi + a_lo)(w_hi + w_lo) without lo * lo
    ah = _bf16(A); al = _bf16(A - ah); wh = _bf16(W); wl = _bf16(W - wh)
    e3 = err(ah.astype(np.float64) @ wh + ah.astype(np.float64) @ wl + al.astype(np.float64) @ wh)
    # scaled messages: a power of two per row brings the maximum near 2^4; weights scaled into [2^12, 2^13)
    sc = 2.0 ** np.round(np.log2(16 / np.abs(A).max(axis=1, keepdims=True)))
    As = (A * sc).astype(np.float32); a16 = _f16(As)
    wsc = 2.0 ** 13 / (2 ** np.ceil(np.log2(np.abs(W).max())))
    Ws = (W * wsc).astype(np.float32); wh16 = _f16(Ws); wl16 = _f16(Ws - wh16)
    # plain two products: one fp16 message plane x (w_hi + w_lo)
    e2 = err((a16.astype(np.float64) @ (wh16.astype(np.float64) + wl16)) / sc / wsc)
    # fp16 + fp8: a16 * w_hi (kind::f16) + [e4m3(a16) | e4m3((v - a16) 2^13)] * [e4m3(w_lo) ; e4m3(w 2^-13)] (kind::f8f6f4)
    r8 = _e4m3((As - a16) * 2.0 ** 13); w8 = _e4m3(Ws / 2.0 ** 13); a8 = _e4m3(a16); wl8 = _e4m3(wl16)
    e8 = err((a16.astype(np.float64) @ wh16.astype(np.float64) + a8.astype(np.float64) @ wl8.astype(np.float64)
              + r8.astype(np.float64) @ w8.astype(np.float64)) / sc / wsc)
    return e3, e2, e8


@pytest.mark.parametrize("K,tail", [(576, 1.5), (2304, 1.5), (1152, 0.0)])
def test_backward_operand_formats(K, tail):
    e3, e2, e8 = _errors(K, tail, 0)
    assert e3 < 1e-5                      # three bf16 products: 16 bits of both operands
    assert 5e-5 < e2 < 5e-4               # one fp16 message plane: 11 bits of the message
    assert e8 < 4e-5 and e8 < e2 / 5      # fp16 + fp8: ~15 bits for two product-equivalents


def test_e4m3_planes_stay_in_range():
    """The byte planes must not saturate: the message's top bits (|a16| <= 2^4 .. 2^6 after scaling, E4M3 max 448) and the
    residual scaled by 2^13 (|v - a16| <= 2^-11 |v| -> <= 2^8 for |v| < 2^6)."""
    v = np.float32(63.9)
    a16 = _f16(np.array([v]))[0]
    assert abs(_e4m3(np.array([a16]))[0] - a16) <= a16 / 16 + 1e-6          # 3 mantissa bits, no saturation
    worst_residual = 2.0 ** -11 * 64 * 2.0 ** 13
    assert worst_residual <= 448 and np.isfinite(_e4m3(np.array([worst_residual], dtype=np.float32))[0])
