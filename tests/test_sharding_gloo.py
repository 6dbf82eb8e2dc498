"""CPU, world_size 2 over gloo: the N>1 path is a pure partition of images over ranks -- each rank explains its own
block and the gathered maps equal the single-process result.  (The per-rank compute is stood in for by the oracle on
tiny dimensions; on the GPU box the same sharding code feeds the CUDA engine, bench.py --gpus N.)"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _explain_block(images, caps):
    from lrp_imagecaptioning_b200 import synth
    from oracle.decoder_ref import DecoderRef
    dec = synth.decoder_weights("adaptive", V=40, H=8, E=8, D=12, seed=3)
    out = []
    for F, cap in zip(images, caps):
        o = DecoderRef(dec).forward(F, list(cap))
        out.append(np.stack([o.explain(t)[0].reshape(-1) for t in range(1, len(cap) + 1)]))
    return np.stack(out) if out else np.zeros((0, caps.shape[1], images.shape[1] * images.shape[2]), np.float32)


def _worker(rank, world, port, n_images, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lrp_imagecaptioning_b200 import synth
    from lrp_imagecaptioning_b200.engine import shard_images
    F = synth.features(n_images, L=4, D=12, seed=5)
    caps = synth.captions(n_images, 3, 40, seed=6)
    mine = shard_images(n_images, rank, world)
    local = torch.from_numpy(_explain_block(F[mine], caps[mine]))
    sizes = [len(shard_images(n_images, r, world)) for r in range(world)]
    bufs = [torch.zeros((s,) + tuple(local.shape[1:]), dtype=local.dtype) for s in sizes]
    dist.all_gather(bufs, local) if len(set(sizes)) == 1 else dist.all_gather_object(bufs, local)
    if rank == 0:
        ret["gathered"] = torch.cat([torch.as_tensor(b) for b in bufs]).numpy()
        ret["single"] = _explain_block(F, caps)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max-over-ranks timing reduction used by bench.py
    assert t.item() == world
    dist.destroy_process_group()


def test_sharded_explanations_equal_single_process():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, 6, ret), nprocs=2, join=True)
    assert ret["gathered"].shape == ret["single"].shape
    assert np.array_equal(ret["gathered"], ret["single"])   # units are independent: bit-for-bit


def test_uneven_shards():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29534, 5, ret), nprocs=2, join=True)
    assert np.array_equal(ret["gathered"], ret["single"])


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lrp_imagecaptioning_b200.lrp_inference import allreduce_mean_
    g = torch.Generator().manual_seed(100 + rank)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
    local = [torch.randn(p.shape, generator=g) for p in params]
    for p, l in zip(params, local):
        p.grad = l.clone()
    flat = allreduce_mean_(params)
    ret["local%d" % rank] = [l.numpy() for l in local]
    ret["reduced%d" % rank] = [p.grad.numpy().copy() for p in params]
    ret["flat%d" % rank] = flat.numpy().copy()
    dist.destroy_process_group()


def test_gradient_allreduce_is_the_mean_of_the_per_rank_gradients():
    """The fine-tune step's one collective (lrp_inference.allreduce_mean_, train.py:569-577 on N GPUs): after it every rank
    holds (sum of the per-rank gradients) / world, parameter by parameter."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, 29535, ret), nprocs=2, join=True)
    for k in range(2):
        want = (ret["local0"][k] + ret["local1"][k]) / 2.0
        for r in range(2):
            assert np.allclose(ret["reduced%d" % r][k], want, rtol=0, atol=1e-7)
    assert np.array_equal(ret["flat0"], ret["flat1"])
    assert np.allclose(ret["flat0"], np.concatenate([((ret["local0"][k] + ret["local1"][k]) / 2.0).reshape(-1) for k in range(2)]), atol=1e-7)
