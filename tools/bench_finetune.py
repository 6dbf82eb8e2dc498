"""Timing of the LRP-inference fine-tune step (BASELINE.json configs[4]; reference train.py:569-577): predict, explain
every predicted non-stop word down to pixels, weight the logits, one Adam step.  One rank per GPU (torchrun) with the
gradient all-reduce over NCCL, or a single process.  Writes gpurun_out/bench_finetune.json on rank 0.
Usage: python tools/bench_finetune.py [--batch 64] [--steps 3] [--kind adaptive]"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrp_imagecaptioning_b200 import synth
from lrp_imagecaptioning_b200.model import CaptioningModel
from lrp_imagecaptioning_b200.lrp_inference import LRPInferenceTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--kind", default="adaptive")
ap.add_argument("--vocab", type=int, default=10000)
ap.add_argument("--breakdown", action="store_true", help="also time the phases of one step (synchronising between them)")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
T, HW = 20, 224


class _Pre(object):
    SOS_TOKEN_LABEL_ENCODED, EOS_TOKEN_LABEL_ENCODED = 1, 2
    _word_of = {}


class _Provider(object):
    caption_preprocessor = _Pre()


model = CaptioningModel.synthetic(args.kind, vocab_size=args.vocab, image_hw=HW, seed=0, precision="tc", device="cuda:%d" % lr)
tr = LRPInferenceTrainer(model, _Provider(), "mean", learning_rate=1e-5, stop_words=set())
g = np.random.default_rng(rank)
imgs_host = torch.from_numpy(synth.images(args.batch, HW, 50 + rank)).pin_memory()
cap = g.integers(3, args.vocab - 1, size=(args.batch, T))
tok_in = np.concatenate([np.ones((args.batch, 1), int), cap[:, :-1]], axis=1)
y = (cap - 1).astype(np.int64)            # class indices (the one-hot (B, T, V) tensor of the Keras loop is never built)


def one_step():
    imgs = imgs_host.to("cuda:%d" % lr, non_blocking=True)     # the step's images arrive from (pinned) host memory
    return tr.step(tok_in, imgs, y)


one_step()              # warm-up (allocations, CUDA graph capture of the decoder forward)
one_step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time(); words = 0
for _ in range(args.steps):
    loss = one_step()
    words += tr.explained_words
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.time() - t0) / args.steps
if world > 1:
    t = torch.tensor([dt], device="cuda:%d" % lr)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
breakdown = None
if args.breakdown:
    import types
    ph = {}

    def timed(name, fn):
        torch.cuda.synchronize(); t = time.time(); r = fn(); torch.cuda.synchronize(); ph[name] = ph.get(name, 0.0) + time.time() - t
        return r
    eng = tr.layer._engine
    imgs = imgs_host.to("cuda:%d" % lr)
    timed("sync_engine (device-to-device weights, re-derived layouts)", lambda: tr.net.sync_engine(eng))
    pred = timed("predict (encoder forward + teacher-forced decoder)", lambda: tr.predict(tok_in, imgs))
    sw = timed("sparse_weights (decoder + encoder relevance of every predicted word)", lambda: tr.layer.sparse_weights(imgs, pred, features_ready=True))
    tok = torch.as_tensor(tok_in, device=imgs.device).long()
    logits = timed("torch forward (differentiable model pass)", lambda: tr.net(tok, imgs))
    timed("torch backward", lambda: logits.float().sum().backward())
    breakdown = {k: round(v, 4) for k, v in ph.items()}
if rank == 0:
    out = {"kind": args.kind, "batch_per_gpu": args.batch, "gpus": world, "vocab": args.vocab, "s_per_step": dt,
           "steps_per_s": 1.0 / dt, "explained_words_per_step_per_gpu": words / args.steps,
           "explained_words_per_s_all_gpus": world * words / args.steps / dt, "loss": loss}
    if breakdown:
        out["phases_s"] = breakdown
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/bench_finetune.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
