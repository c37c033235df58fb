#!/usr/bin/env python
"""Device time of one LA-VAE training step (vqvae.shared_eval 'train': forward + losses + backward, optimizer excluded)
per batch size / length, univariate (vqvae.py) and the fork's multivariate config (myvqvae.py, input_dim 7, flow_dim 50)."""
import json
import os
import sys
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200 import synth
from t2ms_b200.compat import VAE_ARGS
from t2ms_b200.lavae import vqvae as vq_uni
from t2ms_b200.mylavae import vqvae as vq_multi

DEV = "cuda:0"
rows = []
for cin, flow, L, B in ((1, 30, 24, 256), (1, 30, 96, 256), (1, 30, 96, 1024), (7, 50, 100, 256)):
    m = vq_uni(VAE_ARGS) if cin == 1 else vq_multi(Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256,
                                                             embedding_dim=64, flow_dim=flow, input_dim=cin))
    m.load_state_dict(synth.make_vae_state(1, in_channels=cin))
    m = m.to(DEV).train()
    x = torch.rand(B, L, device=DEV) if cin == 1 else torch.rand(B, cin, L, device=DEV)
    eng = m._engine if cin > 1 else None
    if eng is None:
        from t2ms_b200.lavae_train import LavaeEngine
        eng = LavaeEngine(m, 30)
    for _ in range(3):
        eng.step(x, backward=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.step(x, backward=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flop = 3 * 2 * (705_536 + 640_000) * (L // 4) * B if cin == 1 else None      # encoder + decoder MACs x (fwd + 2 bwd)
    rows.append({"in_channels": cin, "flow_dim": flow, "length": L, "batch": B, "ms_per_step": round(ms, 3), "series_per_s": round(B / ms * 1e3, 1),
                 "tflops_fp32_algorithmic": round(flop / ms / 1e9, 2) if flop else None})
print(json.dumps(rows, indent=1))
