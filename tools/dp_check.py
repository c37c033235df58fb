#!/usr/bin/env python
"""Data-parallel check (torchrun, one rank per GPU, NCCL): (1) sharded sampling + final all_gather equals the
single-GPU result; (2) a DitTrainer step on batch shards equals the step on the whole batch (gradient all-reduce),
and every rank ends with identical parameters.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200 import T2SSampler, Transformer, synth, vqvae
from t2ms_b200.compat import VAE_ARGS
from t2ms_b200.sampler import gather_series, shard_range
from t2ms_b200.training import DitTrainer


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}

    def make():
        m = Transformer()
        m.load_state_dict(synth.make_dit_state(15, bias_std=0.02))
        return m.to(dev)

    # ---- sampling: shard + gather vs whole batch on every rank
    vae = vqvae(VAE_ARGS)
    vae.load_state_dict(synth.make_vae_state(1))
    vae = vae.to(dev).eval()
    B = 8 * world
    emb, noise = synth.make_text_embeddings(B).to(dev), synth.make_noise(B).to(dev)
    smp = T2SSampler(make().eval(), vae)
    full = smp.sample(emb, 48, steps=4, noise=noise)
    lo, hi = shard_range(B, rank, world)
    part = smp.sample(emb[lo:hi], 48, steps=4, noise=noise[lo:hi])
    got = gather_series(part, B)
    res["sampling_gather_max_abs"] = (got - full).abs().max().item()

    # ---- training: sharded step vs whole-batch step
    x_t, tgt = torch.randn(B, 64, 30, device=dev, generator=torch.Generator(dev).manual_seed(3)), None
    g = torch.Generator(dev).manual_seed(4)
    tgt = torch.randn(B, 64, 30, device=dev, generator=g)
    t = torch.rand(B, device=dev, generator=g)
    ref = DitTrainer(make().train(), group=dist.new_group([rank]))          # world-1 group: no exchange
    loss_ref = ref.step(x_t, t, emb, tgt)
    dp = DitTrainer(make().train())
    loss_dp = dp.step(x_t[lo:hi], t[lo:hi], emb[lo:hi], tgt[lo:hi])
    torch.cuda.synchronize()
    res["train_loss_rel"] = abs(loss_dp.item() - loss_ref.item()) / loss_ref.item()
    res["train_grad_rel_l2"] = rel(dp.grads.flat, ref.grads.flat)
    res["train_param_update_rel_l2"] = rel(dp.params.flat - ref.params.flat + 1e-30, torch.zeros_like(ref.params.flat) + 1e-30) if False else \
        (dp.params.flat - ref.params.flat).abs().max().item()
    mine = dp.params.flat.clone()
    dist.broadcast(mine, src=0)
    res["params_identical_across_ranks"] = bool(torch.equal(mine, dp.params.flat))
    # ---- the shared CFG-dropout coin (train.py:80-82): all ranks must take the same branch
    torch.manual_seed(100 + rank)                                           # different CPU RNG streams per rank
    flags = []
    for _ in range(6):
        before = dp.loss_sum.clone()
        dp.train_batch(x_t[lo:hi], emb[lo:hi])
        flags.append(0)
    p2 = dp.params.flat.clone()
    dist.broadcast(p2, src=0)
    res["params_identical_after_train_batches"] = bool(torch.equal(p2, dp.params.flat))
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
