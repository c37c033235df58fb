#!/usr/bin/env python
"""Per-kernel device times of one guided step at a batch size (CUDA events around the stage-wise C-ABI entries, L2-warm) next to
the time of the same step inside the enqueue-only loop: what the launch boundaries cost at small batches.
    python tools/step_kernels.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, Workspace, make_dit, stream
from t2ms_b200 import T2SSampler, _lib, synth

lib = _lib.load()
model, _ = make_dit(0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = 2 * batch
x = torch.randn(batch, 64, 30, device=DEV)
emb = torch.randn(batch, 128, device=DEV)
t100 = torch.full((1,), 37.0, device=DEV)
out = torch.empty(B, 64, 30, device=DEV)
pk = model.packed()
ws = Workspace(model, B)
calls = [("cond", lambda: lib.t2s_dit_cond(pk.ref, t100.data_ptr(), 0, emb.data_ptr(), 1, B, ws.ptr, stream())),
         ("EMBED", lambda: lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream())),
         ("attention", lambda: lib.t2s_dit_attention(B, ws.ptr, stream())),
         ("MID", lambda: lib.t2s_dit_block_post(pk.ref, 1, B, ws.ptr, stream())),
         ("FINAL", lambda: lib.t2s_dit_final(pk.ref, out.data_ptr(), B, ws.ptr, stream()))]
for _, fn in calls:
    fn()
torch.cuda.synchronize()
tot = 0.0
for (label, fn), mult in zip(calls, (1, 1, 4, 3, 1)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    tot += mult * us
    print(f"  {label:10s} {us:8.1f} us  x{mult}")
print(f"  sum of the ten launches of a step, back to back per kernel: {tot:.1f} us")
smp = T2SSampler(model)
e = synth.make_text_embeddings(batch, seed=7).to(DEV)
x0 = torch.randn(batch, 64, 30, device=DEV)
for pdl in (0, 1):
    lib.t2s_set_pdl(pdl)
    for _ in range(2):
        smp.sample_latent(e, steps=50, noise=x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    smp.sample_latent(e, steps=200, noise=x0)
    e1.record()
    torch.cuda.synchronize()
    print(f"  guided step inside the loop (PDL {'on' if pdl else 'off'}): {e0.elapsed_time(e1) / 200 * 1e3:.1f} us")
