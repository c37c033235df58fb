#!/usr/bin/env python
"""Guided rectified-flow sampling throughput of the denoiser per latent width: Transformer() (H = 30,
480 tokens) and the fork's Transformer(50) / Transformer(64) (800 / 1024 tokens), latents only (no LA-VAE)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200 import T2SSampler, Transformer, synth

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rows = []
for dim in (30, 50, 64):
    m = Transformer(dim)
    m.load_state_dict(synth.make_dit_state(0, dim=dim))
    smp = T2SSampler(m.to(DEV).eval())
    emb = synth.make_text_embeddings(B).to(DEV)
    noise = synth.make_noise(B, dim=dim).to(DEV)
    for _ in range(2):
        smp.sample_latent(emb, steps=STEPS, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        smp.sample_latent(emb, steps=STEPS, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ntok = 16 * dim
    flop_seq = 2 * (ntok * 128 * (384 + 128 + 256 + 256) + 2 * ntok * ntok * 128) * 4      # 4 blocks: linears + QK^T + PV
    rows.append({"dim": dim, "tokens": ntok, "batch": B, "steps": STEPS, "ms_per_step": round(ms / STEPS, 4),
                 "latents_per_s_at_100_steps": round(B / (ms / STEPS * 100 / 1e3), 1),
                 "tflops_algorithmic": round(2 * B * STEPS * flop_seq / (ms / 1e3) / 1e12, 1)})
    del smp, m
    torch.cuda.empty_cache()
print(json.dumps(rows, indent=1))
