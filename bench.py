#!/usr/bin/env python
"""Benchmark of the hot path: explained words / second (full LRP to pixels).

Default workload (BASELINE.json configs[1], "config2"): grid-TD captioner, 64 synthetic 224x224 images per GPU x 20 greedy
words, LRP-epsilon decoder + VGG16 encoder (LRPEpsilon, eps=0.01) -> one 224x224x3 relevance map per word.
One "step" = one pass of the whole path over that batch.  `value` is timed with the images resident in HBM and no
per-launch instrumentation; `e2e` goes through the C-ABI host-buffer call (pinned host images in, pixel maps out) every
step; kernel times for the roofline come from one separate, event-instrumented step.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --workload config3 [--rule a2b1]          adaptive, V = 10 000, 512 images in total STRONG-scaled over
                                                            the ranks, LRP-alpha-beta (BASELINE.json configs[2])
  python bench.py --workload config1                        one image x 20 words (latency, BASELINE.json configs[0])
  python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference's algorithm)
The default run also appends, as extra keys of the same JSON line: `config3` (the north-star configuration, two timed
steps), `latency` (configs[0] on the GPU), and for N > 1 `gather` (NCCL all-gather of the per-word heat maps) and
`cross_rank_check` (rank r recomputes the first image of rank r+1 and compares the maps bit for bit).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:   # the CPU arm may use every host core (torchrun exports OMP_NUM_THREADS=1)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "explained words/sec (full LRP to pixels)"
UNIT = "words/s"
T_WORDS, VOCAB, HW = 20, 10000, 224
ENC_GFLOP_PER_WORD = 30.69      # one transposed-conv sweep of VGG16
# measured DRAM read + write of the 12 transposed-conv launches per word (ncu --set full, 320 words, tools/profile_encoder.py):
# fp16 + fp8 messages 33.07 GB (profiles/r02_ncu_enc_eps_summary.csv), one fp16 plane 17.80 GB (r02_ncu_enc_preseta_summary.csv),
# two bf16 planes 31.4 GB (r01c_ncu_full_bwd_summary.csv)
NCU_DRAM_MB_PER_WORD = {"h1f8": (103.4, "profiles/r02_ncu_enc_eps_summary.csv", 95.1), "f16x2": (55.6, "profiles/r02_ncu_enc_preseta_summary.csv", 47.6),
                        "bf16x3": (98.2, "profiles/r01c_ncu_full_bwd_summary.csv", 95.1), "fp32": (None, "not captured", None)}

WORKLOADS = {
    # name: (decoder kind, images per GPU (weak) or total (strong), scaling, description)
    "config2": ("gridtd", 64, "weak", "configs[1]: grid-TD, %d images/GPU x 20 greedy words, LRP-eps decoder + VGG16 LRPEpsilon(0.01) encoder, 224x224"),
    "config3": ("adaptive", 512, "strong", "configs[2]: adaptive attention, V=10000, %d images in total x 20 greedy words, LRP-eps decoder + VGG16 LRP-alpha-beta (%s) encoder, 224x224"),
    "config1": ("adaptive", 1, "weak", "configs[0]: adaptive attention, %d image x 20 greedy words, LRP-eps decoder + VGG16 LRPEpsilon(0.01) encoder, 224x224"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(object):
    def __init__(self, index):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """The timed region starts here (the sampler itself started before the warm-up steps: nvidia-smi needs ~0.3 s to come up)."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        window = "timed region"
        rows = self.rows[self.first:]
        if not rows:   # a timed region shorter than the sampling period: the warm-up steps ran the same kernels
            rows, window = list(self.rows), "warm-up + timed region (timed region shorter than the 100 ms sampling period)"
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def rule_for(workload, rule_name):
    from lrp_imagecaptioning_b200 import _lib
    from lrp_imagecaptioning_b200.encoder import RuleSpec
    if workload != "config3":
        return RuleSpec(_lib.RULE_EPSILON, epsilon=0.01, bias=True), "LRPEpsilon(0.01)", 1
    if rule_name == "a2b1":
        return RuleSpec(_lib.RULE_ALPHA_BETA, alpha=2, beta=1, bias=True), "alpha=2, beta=1", 2
    return RuleSpec(_lib.RULE_ALPHA_BETA, alpha=1, beta=0, bias=True), "PresetA: alpha=1, beta=0, bias", 1


def images_per_rank(workload, world):
    kind, n, scaling, _ = WORKLOADS[workload]
    if scaling == "strong":
        if n % world:
            raise SystemExit("bench.py: %d images do not split evenly over %d ranks" % (n, world))
        return n // world
    return n


def config_dict(args, world, workload=None, rule_desc=""):
    workload = workload or args.workload
    kind, n, scaling, desc = WORKLOADS[workload]
    text = desc % ((n, rule_desc) if workload == "config3" else (n,))
    return {"workload": text, "images_per_gpu": images_per_rank(workload, world), "words_per_image": T_WORDS, "vocab": VOCAB,
            "parallelism": "images sharded over the ranks (one rank per GPU), no data-path collective",
            "precision": args.precision, "l2": "inputs larger than L2 (relevance messages are GBs per layer)",
            "chunk_words": getattr(args, "chunk_words", None)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(words, seed=0, kind="gridtd"):
    """The reference's algorithm on host cores for ONE image: faithful oracle port of the NumPy decoder (dense attribution
    matrices, per-cell loops, oracle/decoder_ref.py faithful=True) + torch-CPU restatement of the iNNvestigate epsilon
    rule.  `words`: the 1-based positions to explain (the reference's cost grows with the position: `t x L` helper calls,
    explainers.py:1292-1299).  Returns (seconds for the image forward, [seconds per listed word])."""
    import torch
    from lrp_imagecaptioning_b200 import synth
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    torch.set_num_threads(os.cpu_count() or 1)
    vgg = synth.vgg16_weights(seed)
    dec = synth.decoder_weights(kind, V=VOCAB, seed=seed + 1)
    x = synth.images(1, HW, seed + 2)
    cap = list(synth.captions(1, T_WORDS, VOCAB, seed=seed + 3)[0])
    t0 = time.time()
    F = ER.features(x, vgg)
    o = DecoderRef(dec, faithful=True).forward(F[0].reshape(-1, 512), cap)
    t_img = time.time() - t0
    t_words = []
    for t in words:
        t0 = time.time()
        rF, _ = o.explain(int(t))
        ER.analyze("lrp.epsilon", x, rF, vgg, epsilon=0.01)
        t_words.append(time.time() - t0)
    return t_img, t_words


def stratified_positions(k):
    """k word positions spread evenly over 1..T (k = T: every position once), so that the sample mean of the position
    -- which the reference's per-word cost is proportional to -- matches the workload's (10.5)."""
    return [int(min(T_WORDS, max(1, round(0.5 + (j + 0.5) * T_WORDS / k)))) for j in range(k)]


def run_reference(args, rank):
    if rank != 0:
        return
    pos = stratified_positions(max(args.steps, 1))
    t_imgs, t_words = [], []
    for i in range(args.warmup):
        cpu_reference_sample([pos[i % len(pos)]], seed=1000 + i)
    for i in range(args.steps):
        t_img, tw = cpu_reference_sample([pos[i]], seed=i)
        t_imgs.append(t_img)
        t_words.append(tw[0])
    t_img = float(np.mean(t_imgs))
    t_word = float(np.mean(t_words))
    value = T_WORDS / (t_img + T_WORDS * t_word)
    cores = os.cpu_count() or 1
    sample = "one reference step = 1 image: forward (VGG16 + 20-step decoder) + ONE word of it through decoder-LRP + encoder " \
             "LRP-eps to pixels; the %d timed steps take the word positions %s (stratified over 1..20: the reference's " \
             "per-word cost grows with the position); words/s = 20 / (mean t_image + 20 * mean t_word) -- the reference is " \
             "strictly serial per word" % (args.steps, pos)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * (t_img + t_word), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64 decoder / f32 encoder", "data": "synthetic",
           "config": config_dict(args, 1, "config2"),
           "step_definition": "1 image forward + 1 explained word (ms_per_step); value extrapolates to 20 words per image",
           "t_image_s": t_img, "t_word_s": t_word,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
class Job(object):
    """One workload on this rank: model, engine, pinned host buffers, resident images."""

    def __init__(self, args, workload, rank, world, dev, rule_name="presetA"):
        import torch
        from lrp_imagecaptioning_b200 import synth
        from lrp_imagecaptioning_b200.engine import ExplainEngine, word_list
        from lrp_imagecaptioning_b200.model import CaptioningModel
        self.torch = torch
        self.workload = workload
        self.kind = WORKLOADS[workload][0]
        self.n_img = images_per_rank(workload, world)
        self.rule, self.rule_desc, self.flop_mul = rule_for(workload, rule_name)
        self.model = CaptioningModel.synthetic(self.kind, vocab_size=VOCAB, image_hw=HW, seed=0, precision=args.precision, device=dev)
        self.eng = ExplainEngine(self.model, rule=self.rule)
        self.model.image_model.set_chunk_words(args.chunk_words)
        if args.promote is not None:
            self.model.image_model.set_promote(args.promote)
        self.seed = 100 + rank
        self.x_host = torch.from_numpy(synth.images(self.n_img, HW, self.seed)).pin_memory()
        self.x_dev = self.x_host.to(dev)
        self.wi, self.wt = word_list(self.n_img, T_WORDS)
        self.n_words = len(self.wi)
        self.cap_host = np.zeros((self.n_img, T_WORDS), dtype=np.int32)
        self.out_host = None

    def step_resident(self):
        self.eng.forward(self.x_dev, T=T_WORDS, greedy=True)
        return self.eng.explain_words(self.wi, self.wt)

    def step_e2e(self):
        if self.out_host is None:
            self.out_host = self.torch.empty((self.n_words, HW, HW, 3), dtype=self.torch.float32).pin_memory()
        out_np = self.out_host.numpy()
        self.eng.explain_batch_host(self.x_host.numpy(), self.cap_host, greedy=True, out=out_np)
        return float(out_np[0, 0, 0, 0])

    def close(self):
        self.model.image_model.close()
        self.eng.decoder.close()
        self.out_host = None
        self.x_dev = None
        self.torch.cuda.empty_cache()


def run_ours(args, rank, local_rank, world):
    import torch
    from lrp_imagecaptioning_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = "cuda:%d" % local_rank

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b) / steps
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    job = Job(args, args.workload, rank, world, dev, args.rule)
    im = job.model.image_model
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        job.step_resident()
    torch.cuda.synchronize()
    l0 = job.eng.launches()
    sampler.mark()
    ms = timed(job.step_resident, args.steps, 0)          # no per-launch instrumentation inside this region
    clocks = sampler.stop()
    launches = (job.eng.launches() - l0) // max(args.steps, 1)
    ms_e2e = timed(job.step_e2e, args.steps, max(1, min(args.warmup, 2)))

    # one separate instrumented step: CUDA events around every convolution launch (kernel ms for the roofline) ...
    im.profile(True)
    im.profile_read()
    job.step_resident()
    prof = im.profile_read()
    im.profile(False)
    # ... and the phase breakdown of another one (CUDA events on the launching stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    im.forward(job.x_dev, job.rule)
    feats = im.features()
    ev[1].record()
    job.eng.decoder.forward(feats, T=T_WORDS, greedy=True, eos=job.eng.eos)
    ev[2].record()
    R_head, _, _ = job.eng.decoder.relevance(job.wi, job.wt, want_words=False, want_attention=False)
    ev[3].record()
    maps = im.relevance(job.wi, R_head.view(-1, HW // 16, HW // 16, R_head.shape[-1]))
    ev[4].record()
    torch.cuda.synchronize()
    phases = {k: ev[i].elapsed_time(ev[i + 1]) for i, k in enumerate(("encoder_forward", "decoder_forward", "decoder_relevance", "encoder_relevance"))}

    extra = {}
    if dist is not None and not args.no_extras:
        # (C2) gather of the per-word heat maps (channel mean, 200 KB per word) over NCCL, timed on the device
        from lrp_imagecaptioning_b200 import evaluation
        hm = evaluation.heatmaps(maps, mode="mean", device=dev)
        allhm = torch.empty((world,) + tuple(hm.shape), dtype=hm.dtype, device=dev)
        dist.all_gather_into_tensor(allhm, hm)     # warm-up (communicator set-up)
        g_ms = timed(lambda: dist.all_gather_into_tensor(allhm, hm), 3, 0)
        extra["gather"] = {"what": "NCCL all_gather of the channel-mean heat maps [words, 224, 224] fp32 of every rank",
                           "ms": g_ms, "bytes_per_rank": int(hm.numel() * 4), "bytes_total": int(allhm.numel() * 4),
                           "GBps_per_rank_in": allhm.numel() * 4 * (world - 1) / world / (g_ms / 1e3) / 1e9}
        # cross-rank check: rank r recomputes the first image of rank r+1 (as a batch of one) and compares bit for bit
        nb = (rank + 1) % world
        x_nb = torch.from_numpy(synth.images(job.n_img, HW, 100 + nb)[:1]).to(dev)
        mine = maps[:T_WORDS].contiguous().view(torch.int32).to(torch.int64).sum().reshape(1)
        job.eng.forward(x_nb, T=T_WORDS, greedy=True)
        wi1 = np.zeros(T_WORDS, dtype=np.int32)
        wt1 = np.arange(1, T_WORDS + 1, dtype=np.int32)
        theirs = job.eng.explain_words(wi1, wt1).contiguous().view(torch.int32).to(torch.int64).sum().reshape(1)
        sums = torch.empty((world, 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, mine)
        ok = torch.tensor([1 if int(sums[nb, 0].item()) == int(theirs.item()) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        extra["cross_rank_check"] = {"what": "rank r explains image 0 of rank r+1 alone (20 words) and compares the bit pattern "
                                             "checksum of the 20 maps with the owner's", "bit_identical_on_all_ranks": bool(ok.item())}
    del maps, R_head, feats

    total_words = job.n_words * world
    value = total_words / (ms / 1000.0)
    same_sign = args.workload == "config3"
    mode = {"tc": "f16x2" if same_sign else "h1f8", "bf16x3": "bf16x3", "f16x2": "f16x2", "h1f8": "h1f8", "fp32": "fp32"}[args.precision]
    products = 3 if mode == "bf16x3" else 2
    how = {"bf16x3": "hi*hi + hi*lo + lo*hi, bf16 planes", "f16x2": "a*w_hi + a*w_lo, one scaled fp16 message plane",
           "h1f8": "one kind::f16 product a16*w_hi plus one double-rate kind::f8f6f4 product [a8 | r8]*[w_lo8 ; w8] = two "
                   "product-equivalents", "fp32": "fp32 FMA"}[mode]
    peak_tf, peak_bw, peak_src = peaks()
    tc_ms, tc_flops, tc_n = prof["tc_bwd"]
    achieved = (tc_flops / 1e12) / (tc_ms / 1e3) if tc_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "tc_conv_kernel / tc_conv_vh_kernel <BN, EPI_BWD> (tcgen05 transposed conv + fused rule epilogue)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "traffic": (NCU_DRAM_MB_PER_WORD[mode][0] * 1e6 * job.n_words * job.flop_mul / max(tc_n, 1.0)) if NCU_DRAM_MB_PER_WORD[mode][0] else None,
                "traffic_note": "dram read+write of the 12 transposed-conv launches from ncu --set full (%s, 320 words: %s MB/word; "
                                "algorithmic: %s MB/word of messages + the per-image multipliers), scaled to this run's words per launch"
                                % (NCU_DRAM_MB_PER_WORD[mode][1], NCU_DRAM_MB_PER_WORD[mode][0], NCU_DRAM_MB_PER_WORD[mode][2]),
                "peak_source": "%s bf16 cuBLAS (sustained)" % peak_src,
                "note": "achieved = algorithmic fp32-equivalent FLOPs (2*MAC of the transposed convs, %.2f GFLOP/word) / CUDA-event "
                        "kernel time of one instrumented step; every algorithmic MAC costs %d 16-bit tensor-core MAC-equivalents (%s), "
                        "so tensor-pipe work is %dx this figure" % (ENC_GFLOP_PER_WORD * job.flop_mul, products, how, products),
                "backward_arithmetic": mode,
                "products_per_mac": products,
                "tensor_pipe_frac": products * achieved / peak_tf if peak_tf else None,
                "tensor_pipe_frac_note": "product-equivalents / measured cuBLAS bf16 peak; it can pass 1.0: the sustained cuBLAS figure is "
                                         "itself taken under the power cap, and in the h1f8 mode half of the product-equivalents are "
                                         "kind::f8f6f4 MMAs that run at twice the bf16 rate",
                "kernel_ms_per_step": tc_ms, "kernel_launches_per_step": tc_n,
                "share_of_step": tc_ms / ms if ms > 0 else None,
                "forward_tc_ms_per_step": prof["tc_fwd"][0], "last_dgrad_ms_per_step": prof["last"][0]}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": WORKLOADS[args.workload][2], "vs_baseline": None,
           "dtype": "f32 encoder / f64 decoder" if args.precision == "fp32" else
                    "f32 emulated on the 16-/8-bit tensor pipe (backward: %s = %d product-equivalents per MAC; forward: 3 products on f16 planes), fp32 accumulate; f64 decoder rules" % (mode, products),
           "data": "synthetic", "config": config_dict(args, world, args.workload, job.rule_desc), "clocks": clocks,
           "gpu_launches": int(launches),
           "e2e": {"value": total_words / (ms_e2e / 1000.0), "unit": UNIT, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(job.x_host.numel() * 4),
                   "d2h_bytes_per_step": int(job.n_words * HW * HW * 3 * 4 + job.cap_host.nbytes)},
           "phases_ms": phases, "roofline": roofline}
    out.update(extra)
    job.close()

    if args.workload == "config2" and not args.no_extras:
        # the north-star configuration, strong-scaled over the ranks (two timed steps after one warm-up)
        j3 = Job(args, "config3", rank, world, dev, "presetA")
        ms3 = timed(j3.step_resident, 2, 1)
        out["config3"] = {"workload": config_dict(args, world, "config3", j3.rule_desc)["workload"], "scaling": "strong",
                          "images_per_gpu": j3.n_img, "value": j3.n_words * world / (ms3 / 1e3), "unit": UNIT, "ms_per_step": ms3,
                          "steps": 2, "warmup": 1}
        j3.close()
        if world == 1:
            j1 = Job(args, "config1", rank, world, dev)
            ms1 = timed(j1.step_e2e, 5, 2)
            out["latency"] = {"workload": config_dict(args, 1, "config1")["workload"], "ms": ms1,
                              "what": "one image, 20 words, host buffers in and out (lrpcap_explain_batch_host)",
                              "words_per_s": T_WORDS / (ms1 / 1e3)}
            j1.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "config2":
        pos = [1, 10, 20]
        t_img, tw = cpu_reference_sample(pos)
        t_word = float(np.mean(tw))
        out["cpu_baseline"] = {"value": T_WORDS / (t_img + T_WORDS * t_word), "unit": UNIT, "cores": os.cpu_count() or 1,
                               "kind": "port",
                               "sample": "1 image: VGG16 + 20-step decoder forward (%.1f s) and 3 of its 20 words (positions 1, 10, 20) "
                                         "through decoder-LRP + encoder LRP-eps to pixels (%.2f / %.2f / %.2f s); words/s = "
                                         "20 / (t_image + 20 mean t_word)" % (t_img, tw[0], tw[1], tw[2])}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--rule", default="presetA", choices=["presetA", "a2b1"], help="config3 only: the alpha-beta rule")
    ap.add_argument("--precision", default="tc", choices=["tc", "bf16x3", "f16x2", "h1f8", "fp32"],
                    help="tc: tensor cores; fp16 two-product backward for alpha1-beta0 / z+, fp16 + fp8 for the other rules")
    ap.add_argument("--chunk-words", type=int, default=320)
    ap.add_argument("--promote", type=int, default=None, help="backward accumulator promotion interval (k-steps; 0 = off, -1 = default policy)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config3 / latency / gather / cross-rank extras")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
