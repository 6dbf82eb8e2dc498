"""GPU parity on the BENCHMARKED configurations themselves, end to end (BASELINE.json configs[1] and configs[2]):
224 x 224 images, V = 10 000, T = 20 greedy words, every word of every image: images -> VGG16 -> decoder forward ->
decoder relevance -> encoder relevance -> pixels, against the oracle pipeline the reference's explain_image.py:45-87 runs
(models/explainers.py:1092-1321 / 370-666 restated in oracle/decoder_ref.py and pinned to the reference's own code,
innvestigate rules restated in oracle/encoder_ref.py).  The encoder oracle is pinned to the CUDA forward's max-pool
arg-max routes (tests/test_gpu_encoder.py explains why); tolerances are the north-star ones, per word, no medians.
"""
import numpy as np
import pytest

from tests.util import linf_rel, l2_rel, record, same_topk, sum_err, topk_cells, topk_features

pytestmark = pytest.mark.gpu

HW, V, T = 224, 10000, 20


def _run_config(kind, rule_spec, oracle_method, oracle_kw, n_images, what):
    import torch
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.engine import ExplainEngine, word_list
    from lrp_imagecaptioning_b200.model import CaptioningModel
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    model = CaptioningModel.synthetic(kind, vocab_size=V, image_hw=HW, seed=0, precision="tc")   # bench.py model + precision
    eng = ExplainEngine(model, rule=rule_spec)
    x = synth.images(n_images, HW, 100)                                                                # bench.py's images
    cap = eng.forward(torch.from_numpy(x).cuda(), T=T, greedy=True)
    assert cap.shape == (n_images, T) and not np.any(cap == eng.eos)
    wi, wt = word_list(n_images, T)
    maps, R_head, rw, att = eng.explain_words(wi, wt, want_side=True)
    maps, R_head = maps.cpu().numpy(), R_head.cpu().numpy()
    routes = model.image_model.pool_routes()
    own = ER.pool_routes(x, model.vgg)
    flips = [int(sum((routes[l][n] != own[l][n]).sum() for l in routes)) for n in range(n_images)]
    F = ER.features(x, model.vgg)
    worst = {"dec_linf": 0.0, "linf": 0.0, "l2": 0.0, "sum": 0.0}
    for n in range(n_images):
        o = DecoderRef(model.dec).forward(F[n].reshape(-1, 512), list(cap[n]))
        lg = o.logits.copy()
        lg[:, eng.eos - 1] = -np.inf
        assert list(np.argmax(lg, axis=1) + 1) == list(cap[n]), "greedy caption differs from the oracle's arg-max"
        rF = np.concatenate([o.explain(t)[0] for t in range(1, T + 1)], axis=0)                        # [T, 14, 14, 512]
        force = ER.Forced({l: r[n:n + 1].repeat(T, axis=0) for l, r in routes.items()})
        ref = ER.analyze(oracle_method, np.repeat(x[n:n + 1], T, axis=0), rF, model.vgg, force=force, **oracle_kw)
        for t in range(1, T + 1):
            w = n * T + t - 1
            got_head = R_head[w].reshape(14, 14, 512)
            md = record("%s decoder R_F img %d t %d" % (what, n, t), got_head, rF[t - 1])
            assert md["linf_rel"] <= 1e-3 and md["l2_rel"] <= 1e-3, (what, n, t, md)
            assert same_topk(got_head, rF[t - 1], 10, sums=lambda a: np.asarray(a, dtype=np.float64).sum(axis=-1).reshape(-1))
            m = record("%s pixels img %d t %d" % (what, n, t), maps[w], ref[t - 1], flips=flips[n])
            assert m["linf_rel"] <= 1e-3 and m["l2_rel"] <= 1e-3, (what, n, t, m)
            assert m["sum_err"] <= 1e-4, (what, n, t, m)
            assert same_topk(maps[w], ref[t - 1], 10), (what, n, t, topk_cells(maps[w], 10), topk_cells(ref[t - 1], 10))
            worst["dec_linf"] = max(worst["dec_linf"], md["linf_rel"])
            worst["linf"] = max(worst["linf"], m["linf_rel"])
            worst["l2"] = max(worst["l2"], m["l2_rel"])
            worst["sum"] = max(worst["sum"], m["sum_err"])
    print("%s: %d words, route flips per image %s, worst %s" % (what, n_images * T, flips, worst))
    return worst


def test_config1_gridtd_lrp_eps_every_word():
    """configs[1]: grid-TD captioner, LRP-eps decoder + LRPEpsilon(0.01) encoder (the driver's bench.py workload)."""
    from lrp_imagecaptioning_b200 import _lib
    from lrp_imagecaptioning_b200.encoder import RuleSpec
    _run_config("gridtd", RuleSpec(_lib.RULE_EPSILON, epsilon=0.01, bias=True), "lrp.epsilon", dict(epsilon=0.01), 2, "config1")


@pytest.mark.parametrize("name,alpha,beta,method", [("presetA", 1, 0, "lrp.sequential_preset_a"), ("a2b1", 2, 1, "lrp.alpha_2_beta_1")])
def test_config3_adaptive_alpha_beta_every_word(name, alpha, beta, method):
    """configs[2] (SURVEY config 3): adaptive-attention captioner, V = 10 000, LRP-alpha-beta encoder (PresetA = alpha1
    beta0 with bias, the reference explainers' default analyzer, explainers.py:32; and alpha2 beta1)."""
    from lrp_imagecaptioning_b200 import _lib
    from lrp_imagecaptioning_b200.encoder import RuleSpec
    _run_config("adaptive", RuleSpec(_lib.RULE_ALPHA_BETA, alpha=alpha, beta=beta, bias=True), method, {}, 2, "config3 " + name)
