"""GPU parity: CUDA decoder forward + word-level relevance (through the C ABI) vs oracle/decoder_ref.py,
which is itself pinned against the reference's own NumPy code (tests/test_oracle_pinning.py, tests/golden)."""
import numpy as np
import pytest

from tests.util import assert_parity, topk_features

pytestmark = pytest.mark.gpu

SMALL = dict(V=60, H=64, E=64, D=96, L=16, T=6, N=3)     # D % 64 != 0: the unfused fp64 / per-GEMM forward
FUSED = dict(V=60, H=64, E=64, D=64, L=16, T=6, N=3)     # every dimension a multiple of 64: the fused forward (decoder_fused.cuh)
FULL = dict(V=10000, H=512, E=512, D=512, L=196, T=20, N=2)   # BASELINE.json sizes: V = 10 000, 20 words


def _setup(kind, cfg, seed=11):
    from lrp_imagecaptioning_b200 import synth
    dec = synth.decoder_weights(kind, V=cfg["V"], H=cfg["H"], E=cfg["E"], D=cfg["D"], seed=seed)
    F = synth.features(cfg["N"], L=cfg["L"], D=cfg["D"], seed=seed + 1)
    cap = synth.captions(cfg["N"], cfg["T"], cfg["V"], seed=seed + 2)
    return dec, F, cap


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
@pytest.mark.parametrize("size", ["small", "fused", "full"])
def test_decoder_lrp_matches_oracle(kind, size):
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    from oracle.decoder_ref import DecoderRef
    cfg = {"small": SMALL, "fused": FUSED, "full": FULL}[size]
    dec, F, cap = _setup(kind, cfg)
    eng = DecoderEngine(dec)
    eng.forward(F, cap)
    N, T = cfg["N"], cfg["T"]
    wi = np.repeat(np.arange(N), T).astype(np.int32)
    wt = np.tile(np.arange(1, T + 1), N).astype(np.int32)
    R, rw, att = eng.relevance(wi, wt)
    R = R.cpu().numpy()
    lk = eng.caption_logits()
    for n in range(N):
        o = DecoderRef(dec).forward(F[n], list(cap[n]))
        ref_lk = np.array([o.logits[i, cap[n, i] - 1] for i in range(T)])
        assert_parity(lk[n], ref_lk, "%s %s logit_k img %d" % (kind, size, n), rel_tol=1e-5, sum_tol=None)
        for t in range(1, T + 1):
            w = n * T + t - 1
            rF, a = o.explain(t)
            assert_parity(R[w], rF.reshape(cfg["L"], cfg["D"]), "%s %s R_F img %d t %d" % (kind, size, n, t),
                          kind=kind, size=size)
            assert topk_features(R[w], 5) == topk_features(rF.reshape(cfg["L"], cfg["D"]), 5)
            assert_parity(att[w], a, "%s %s attention img %d t %d" % (kind, size, n, t), rel_tol=1e-5, sum_tol=None)
            ref_rw = np.zeros(T)
            ref_rw[:len(o.r_words)] = o.r_words
            if np.abs(ref_rw).max() > 0:
                assert_parity(rw[w], ref_rw, "%s %s r_words img %d t %d" % (kind, size, n, t), sum_tol=None)


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
@pytest.mark.parametrize("size", ["small", "full"])
def test_decoder_gradient_matches_oracle(kind, size):
    """_lstm_decoder_backward (manual BPTT, frozen attention; explainers.py:780-832, 1452-1532)."""
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    from oracle.decoder_ref import DecoderRef
    cfg = SMALL if size == "small" else FULL
    if kind == "adaptive" and cfg["E"] != cfg["H"]:
        pytest.skip("the reference's adaptive gradient assumes E == H")
    dec, F, cap = _setup(kind, cfg, seed=31)
    eng = DecoderEngine(dec)
    eng.forward(F, cap)
    N, T = cfg["N"], cfg["T"]
    wi = np.repeat(np.arange(N), T).astype(np.int32)
    wt = np.tile(np.arange(1, T + 1), N).astype(np.int32)
    R, rw = eng.backward(wi, wt)
    R = R.cpu().numpy()
    for n in range(N):
        o = DecoderRef(dec).forward(F[n], list(cap[n]))
        for t in range(1, T + 1):
            w = n * T + t - 1
            g = o.backward(t)
            assert_parity(R[w], g.reshape(cfg["L"], cfg["D"]), "%s %s grad d_F img %d t %d" % (kind, size, n, t),
                          sum_tol=None, kind=kind, size=size)
            ref_rw = np.zeros(T)
            ref_rw[:len(o.r_words)] = o.r_words
            assert_parity(rw[w], ref_rw, "%s %s grad r_words img %d t %d" % (kind, size, n, t), sum_tol=None)


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
@pytest.mark.parametrize("cfgname", ["small", "fused"])
def test_greedy_caption_is_oracle_argmax(kind, cfgname):
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    from oracle.decoder_ref import DecoderRef
    SMALL = {"small": globals()["SMALL"], "fused": FUSED}[cfgname]
    dec, F, _ = _setup(kind, SMALL, seed=21)
    eng = DecoderEngine(dec)
    cap = eng.forward(F, T=SMALL["T"], greedy=True, eos=2)
    assert cap.shape == (SMALL["N"], SMALL["T"]) and cap.min() >= 1 and not np.any(cap == 2)
    for n in range(SMALL["N"]):
        o = DecoderRef(dec).forward(F[n], list(cap[n]))
        lg = o.logits.copy()
        lg[:, 1] = -np.inf   # EOS (id 2 -> index 1) suppressed
        assert list(np.argmax(lg, axis=1) + 1) == list(cap[n])


def test_out_of_range_word_raises():
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    dec, F, cap = _setup("adaptive", SMALL)
    eng = DecoderEngine(dec)
    eng.forward(F, cap)
    with pytest.raises(NotImplementedError):
        eng.relevance([0], [SMALL["T"] + 1])


def test_ragged_word_list_and_order_independence():
    """Words may come in any order, with repeats and different positions per image."""
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    dec, F, cap = _setup("gridtd", SMALL)
    eng = DecoderEngine(dec)
    eng.forward(F, cap)
    wi = np.array([2, 0, 0, 1, 2, 0], dtype=np.int32)
    wt = np.array([1, 6, 3, 2, 5, 3], dtype=np.int32)
    R, rw, att = eng.relevance(wi, wt)
    R = R.cpu().numpy()
    assert np.array_equal(R[2], R[5])
    perm = np.array([5, 3, 1, 0, 4, 2])
    R2, _, _ = eng.relevance(wi[perm], wt[perm])
    assert np.array_equal(R2.cpu().numpy(), R[perm])


@pytest.mark.parametrize("kind", ["adaptive", "gridtd"])
@pytest.mark.parametrize("cfgname", ["small", "fused"])
def test_greedy_forward_graph_replay_is_identical(kind, cfgname, monkeypatch):
    """The greedy forward is captured into a CUDA graph on a repeated configuration and replayed afterwards: eager,
    captured and replayed calls must give the same captions and the same relevance, also for new features, and the
    same as a decoder that never uses a graph."""
    import torch
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.decoder import DecoderEngine
    SMALL = {"small": globals()["SMALL"], "fused": FUSED}[cfgname]
    dec, F, _ = _setup(kind, SMALL, seed=31)
    F2 = synth.features(SMALL["N"], L=SMALL["L"], D=SMALL["D"], seed=99)
    wi = np.array([0, 1, 2, 2], dtype=np.int32)
    wt = np.array([3, 6, 1, 4], dtype=np.int32)

    def run(eng, feats):
        cap = eng.forward(feats, T=SMALL["T"], greedy=True, eos=2)
        R, rw, _ = eng.relevance(wi, wt)
        return cap.copy(), R.cpu().numpy(), np.asarray(rw).copy()
    monkeypatch.setenv("LRPCAP_DECODER_GRAPH", "0")
    plain = DecoderEngine(dec)
    ref1, ref2 = run(plain, F), run(plain, F2)
    monkeypatch.delenv("LRPCAP_DECODER_GRAPH")
    eng = DecoderEngine(dec)
    launches = []
    for i, feats in enumerate([F, F, F, F, F2, F]):      # eager, eager (key seen), capture, replay, replay, replay
        l0 = eng.launches()
        got = run(eng, feats)
        launches.append(eng.launches() - l0)
        want = ref2 if feats is F2 else ref1
        assert np.array_equal(got[0], want[0]), i
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), i
    assert len(set(launches)) == 1                        # replayed launches are counted like eager ones
