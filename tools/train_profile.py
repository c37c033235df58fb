#!/usr/bin/env python
"""Kernel-level time table of one fused training step (torch.profiler / CUPTI)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

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
for _ in range(2):
    tr.step(x_t, tt, e, tgt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(x_t, tt, e, tgt)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name[:90]
        c = rows.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in rows.values())
print(f"B={B}: total kernel time {tot / 1e3:.3f} ms")
for k, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  x{n:4d}  {k}")
if len(sys.argv) > 2:
    print("-- GEMM launches in issue order (us)")
    evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    i = 0
    for ev in evs:
        if "gemm_tf32" in ev.name:
            print(f"{i:3d} {ev.device_time_total:9.1f}")
            i += 1
