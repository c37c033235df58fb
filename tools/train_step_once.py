#!/usr/bin/env python
"""Three fused DiT training steps (forward + MSE + backward + AdamW) at B latents: the short command ncu wraps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200 import Transformer, synth
from t2ms_b200.training import DitTrainer

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
DIM = int(sys.argv[2]) if len(sys.argv) > 2 else 30          # latent width: 30 (T2S), 50 / 64 (the fork's Transformer(dim))
m = Transformer(DIM)
m.load_state_dict(synth.make_dit_state(15, bias_std=0.02, dim=DIM))
tr = DitTrainer(m.to(DEV).train())
x_t, tgt, tt = torch.randn(B, 64, DIM, device=DEV), torch.randn(B, 64, DIM, device=DEV), torch.rand(B, device=DEV)
e = torch.nn.functional.normalize(torch.randn(B, 128, device=DEV), dim=-1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    if i == 2:
        e0.record()
    tr.step(x_t, tt, e, tgt)
e1.record()
torch.cuda.synchronize()
print(f"B={B} dim={DIM}: third step {e0.elapsed_time(e1):.3f} ms, loss {tr.loss_sum.item() / (B * 64 * DIM):.4f}")
