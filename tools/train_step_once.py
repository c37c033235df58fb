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
m = Transformer()
m.load_state_dict(synth.make_dit_state(15, bias_std=0.02))
tr = DitTrainer(m.to(DEV).train())
x_t, tgt, tt = torch.randn(B, 64, 30, device=DEV), torch.randn(B, 64, 30, device=DEV), torch.rand(B, device=DEV)
e = torch.nn.functional.normalize(torch.randn(B, 128, device=DEV), dim=-1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    if i == 2:
        e0.record()
    tr.step(x_t, tt, e, tgt)
e1.record()
torch.cuda.synchronize()
print(f"B={B}: third step {e0.elapsed_time(e1):.3f} ms, loss {tr.loss_sum.item() / (B * 1920):.4f}")
