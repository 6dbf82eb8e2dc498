#!/usr/bin/env python
"""Benchmark of the hot path: explained words / second (full LRP to pixels).

Workload (BASELINE.json configs[1]): grid-TD captioner, 64 synthetic 224x224 images per GPU x 20 greedy words,
LRP-epsilon decoder + VGG16 encoder (LRPEpsilon, eps=0.01) -> one 224x224x3 relevance map per word.
One "step" = one pass of the whole path over that batch.  `value` is timed with the images resident in HBM;
`e2e` goes through the C-ABI host-buffer call (pinned host images in, pixel maps out) every step.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      (CPU arm: the oracle port of the reference's algorithm)
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
N_IMG, T_WORDS, VOCAB, HW = 64, 20, 10000, 224
ENC_GFLOP_PER_WORD = 30.69      # one transposed-conv sweep of VGG16
NCU_DRAM_MB_PER_WORD = 98.2     # measured: profiles/r01c_ncu_full_bwd_summary.csv (31.4 GB over 320 words)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(object):
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def config_dict(args, world):
    return {"workload": "configs[1]: grid-TD, %d images/GPU x %d greedy words, LRP-eps decoder + VGG16 LRPEpsilon(0.01) encoder, 224x224"
                        % (N_IMG, T_WORDS),
            "images_per_gpu": N_IMG, "words_per_image": T_WORDS, "vocab": VOCAB, "parallelism": "images sharded over %d GPU(s), no data-path collective" % world,
            "precision": args.precision, "l2": "inputs larger than L2 (relevance messages are GBs per layer)",
            "lanes": getattr(args, "lanes", 1), "chunk_words": getattr(args, "chunk_words", None)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(n_words=2, seed=0):
    """The reference's algorithm on host cores: faithful oracle port of the NumPy decoder (dense attribution matrices,
    per-cell loops, oracle/decoder_ref.py faithful=True) + torch-CPU restatement of the iNNvestigate epsilon rule."""
    import torch
    from lrp_imagecaptioning_b200 import synth
    from oracle import encoder_ref as ER
    from oracle.decoder_ref import DecoderRef
    torch.set_num_threads(os.cpu_count() or 1)
    vgg = synth.vgg16_weights(seed)
    dec = synth.decoder_weights("gridtd", V=VOCAB, seed=seed + 1)
    x = synth.images(1, HW, seed + 2)
    cap = list(synth.captions(1, T_WORDS, VOCAB, seed=seed + 3)[0])
    t0 = time.time()
    F = ER.features(x, vgg)
    o = DecoderRef(dec, faithful=True).forward(F[0].reshape(-1, 512), cap)
    t_img = time.time() - t0
    t0 = time.time()
    for t in np.linspace(1, T_WORDS, n_words).astype(int):
        rF, _ = o.explain(int(t))
        ER.analyze("lrp.epsilon", x, rF, vgg, epsilon=0.01)
    t_word = (time.time() - t0) / n_words
    return t_img, t_word


def run_reference(args, rank):
    if rank != 0:
        return
    times = []
    for i in range(args.warmup + args.steps):
        t_img, t_word = cpu_reference_sample(n_words=1, seed=i)
        if i >= args.warmup:
            times.append((t_img, t_word))
    t_img = float(np.mean([a for a, _ in times]))
    t_word = float(np.mean([b for _, b in times]))
    value = T_WORDS / (t_img + T_WORDS * t_word)
    cores = os.cpu_count() or 1
    sample = "per step: 1 image forward (VGG16 + 20-step decoder) + 1 word decoder-LRP + encoder-LRP to pixels; " \
             "words/s = 20 / (t_image + 20 * t_word) (the reference is strictly serial per word)"
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * (t_img + t_word), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64 decoder / f32 encoder", "data": "synthetic",
           "config": config_dict(args, 1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    from lrp_imagecaptioning_b200 import _lib, synth
    from lrp_imagecaptioning_b200.encoder import RuleSpec
    from lrp_imagecaptioning_b200.engine import ExplainEngine, StreamedEngine, word_list
    from lrp_imagecaptioning_b200.model import CaptioningModel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = "cuda:%d" % local_rank
    model = CaptioningModel.synthetic("gridtd", vocab_size=VOCAB, image_hw=HW, seed=0, precision=args.precision, device=dev)
    eng = ExplainEngine(model, rule=RuleSpec(_lib.RULE_EPSILON, epsilon=0.01, bias=True))
    model.image_model.set_chunk_words(args.chunk_words)
    if args.promote is not None:
        model.image_model.set_promote(args.promote)
    x_host = torch.from_numpy(synth.images(N_IMG, HW, 100 + rank)).pin_memory()
    x_dev = x_host.to(dev)
    wi, wt = word_list(N_IMG, T_WORDS)
    n_words = len(wi)

    lanes = max(1, args.lanes)
    seng = StreamedEngine(model, rule=eng.rule, lanes=lanes, chunk_words=args.chunk_words) if lanes > 1 else None
    engines = [l["engine"] for l in seng.lanes] if seng else [eng]

    def step_resident():
        if seng:
            return seng.explain_batch(x_dev, T_WORDS, greedy=True)[0]
        eng.forward(x_dev, T=T_WORDS, greedy=True)
        return eng.explain_words(wi, wt)

    cap_host = np.zeros((N_IMG, T_WORDS), dtype=np.int32)
    out_host = torch.empty((n_words, HW, HW, 3), dtype=torch.float32).pin_memory()
    out_np = out_host.numpy()
    x_np = x_host.numpy()

    def step_e2e():
        (seng or eng).explain_batch_host(x_np, cap_host, greedy=True, out=out_np)
        return float(out_np[0, 0, 0, 0])

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

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    l0 = sum(e.launches() for e in engines)
    for e in engines:
        e.image_model.profile(True)
        e.image_model.profile_read()
    sampler.start()
    ms = timed(step_resident, args.steps, 0)
    clocks = sampler.stop()
    prof = None
    for e in engines:   # kernel times add up over the lanes (their launches queue behind one another on the SMs)
        p = e.image_model.profile_read()
        e.image_model.profile(False)
        prof = p if prof is None else {k: tuple(a + b for a, b in zip(prof[k], p[k])) for k in p}
    launches = (sum(e.launches() for e in engines) - l0) // max(args.steps, 1)
    ms_e2e = timed(step_e2e, args.steps, max(1, min(args.warmup, 2)))

    # phase breakdown of one extra resident step (CUDA events on the launching stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    model.image_model.forward(x_dev, eng.rule)
    feats = model.image_model.features()
    ev[1].record()
    eng.decoder.forward(feats, T=T_WORDS, greedy=True, eos=eng.eos)
    ev[2].record()
    R_head, _, _ = eng.decoder.relevance(wi, wt, want_words=False, want_attention=False)
    ev[3].record()
    model.image_model.relevance(wi, R_head.view(-1, HW // 16, HW // 16, R_head.shape[-1]))
    ev[4].record()
    torch.cuda.synchronize()
    phases = {k: ev[i].elapsed_time(ev[i + 1]) for i, k in enumerate(("encoder_forward", "decoder_forward", "decoder_relevance", "encoder_relevance"))}

    total_words = n_words * world
    value = total_words / (ms / 1000.0)
    peak_tf, peak_bw, peak_src = peaks()
    tc_ms, tc_flops, tc_n = prof["tc_bwd"]
    achieved = (tc_flops / 1e12) / (tc_ms / 1e3) if tc_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "tc_conv_kernel / tc_conv_vh_kernel <BN, EPI_BWD> (tcgen05 transposed conv + fused rule epilogue)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
                "traffic": NCU_DRAM_MB_PER_WORD * 1e6 * n_words / max(tc_n / max(args.steps, 1), 1.0),
                "traffic_note": "dram read+write of the 12 transposed-conv launches from ncu --set full (profiles/r01c_ncu_full_bwd_summary.csv, "
                                "320 words: 31.4 GB = %.1f MB/word; algorithmic: 95.1 MB/word of messages + the per-image multipliers), "
                                "scaled to this run's words per launch" % NCU_DRAM_MB_PER_WORD,
                "peak_source": "%s bf16 cuBLAS (sustained)" % peak_src,
                "note": "achieved = algorithmic fp32-equivalent FLOPs (2*MAC of the transposed convs, %.2f GFLOP/word) / CUDA-event "
                        "kernel time; every algorithmic MAC is 3 bf16 tensor-core MACs (hi*hi + hi*lo + lo*hi), so tensor-pipe "
                        "work is 3x this figure" % ENC_GFLOP_PER_WORD,
                "tensor_pipe_frac": 3.0 * achieved / peak_tf if peak_tf else None,
                "kernel_ms_per_step": tc_ms / max(args.steps, 1), "kernel_launches_per_step": tc_n / max(args.steps, 1),
                "share_of_step": (tc_ms / max(args.steps, 1)) / ms if ms > 0 else None,
                "forward_tc_ms_per_step": prof["tc_fwd"][0] / max(args.steps, 1),
                "last_dgrad_ms_per_step": prof["last"][0] / max(args.steps, 1)}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 as split 16-bit tensor-core operands (3 products: bf16 planes backward, f16 planes forward), fp32 accumulate; f64 decoder" if args.precision == "bf16x3"
                    else "f32 encoder / f64 decoder",
           "data": "synthetic", "config": config_dict(args, world), "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": total_words / (ms_e2e / 1000.0), "unit": UNIT, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4 + cap_host.nbytes)},
           "phases_ms": phases, "roofline": roofline}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t_img, t_word = cpu_reference_sample(n_words=2)
        out["cpu_baseline"] = {"value": T_WORDS / (t_img + T_WORDS * t_word), "unit": UNIT, "cores": os.cpu_count() or 1,
                               "kind": "port",
                               "sample": "1 image: VGG16 + 20-step decoder forward (%.1f s) and 2 of its 20 words through "
                                         "decoder-LRP + encoder LRP-eps to pixels (%.1f s/word); words/s = 20 / (t_image + 20 t_word)"
                                         % (t_img, t_word)}
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
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp32"])
    ap.add_argument("--chunk-words", type=int, default=320)
    ap.add_argument("--lanes", type=int, default=1, help="independent stream/thread lanes over blocks of images")
    ap.add_argument("--promote", type=int, default=None, help="backward accumulator promotion interval (k-steps; 0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
