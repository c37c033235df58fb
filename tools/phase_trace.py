#!/usr/bin/env python
"""Phase timeline of the token-block kernel (clock64 stamps of tile 0 / row 0 of every CTA)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from t2ms_b200 import _lib, synth
from gpu_util import DEV, Workspace, make_dit, stream

lib = _lib.load()
model, sd = make_dit(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.randn(B // 2, 64, 30, device=DEV)
emb = torch.randn(B // 2, 128, device=DEV)
t100 = torch.full((1,), 37.0, device=DEV)
pk = model.packed()
ws = Workspace(model, B)
grid = min((B // 2) * 4, 148)
lib.t2s_dit_cond(pk.ref, t100.data_ptr(), 0, emb.data_ptr(), 1, B, ws.ptr, stream())
lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream())
lib.t2s_dit_attention(B, ws.ptr, stream())
names = {0: "start(epi)", 1: "h loaded", 2: "vec ready", 3: "acc proj", 4: "E0 done", 5: "acc fc1a", 6: "E1a done", 7: "acc fc1b", 8: "E1b done",
         9: "acc fc2", 10: "E2 resid", 11: "h store+LN", 12: "acc q", 13: "q stored", 14: "acc k", 15: "k stored", 16: "acc v", 17: "v stored",
         18: "item done"}
for label, fn in (("MID", lambda: lib.t2s_dit_block_post(pk.ref, 1, B, ws.ptr, stream())),
                  ("EMBED", lambda: lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream()))):
    fn(); torch.cuda.synchronize()
    buf = torch.zeros(grid, 32, dtype=torch.int64, device=DEV)
    lib.t2s_debug_set_phase_trace(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    lib.t2s_debug_set_phase_trace(None)
    b = buf.cpu().double()
    base = b[:, 0:1]
    rel = (b - base)
    print(f"== {label}: last work item of each CTA, mean cycles since item start (over {grid} CTAs), delta from previous stamp")
    prev = 0.0
    for i in [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18]:
        col = rel[:, i]
        m = col[b[:, i] > 0]
        if len(m) == 0:
            continue
        v = m.mean().item()
        print(f"  {names[i]:16s} {v:9.0f}  (+{v - prev:7.0f})")
        prev = v

# ---- attention kernel: softmax warpgroup 0, row 0
lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream())
lib.t2s_dit_attention(B, ws.ptr, stream()); torch.cuda.synchronize()
buf = torch.zeros(B * 4, 32, dtype=torch.int64, device=DEV)
lib.t2s_debug_set_phase_trace(buf.data_ptr())
lib.t2s_dit_attention(B, ws.ptr, stream()); torch.cuda.synchronize()
lib.t2s_debug_set_phase_trace(None)
b = buf.cpu().double()
rel = b - b[:, 0:1]
print(f"== ATTENTION (row 0): mean cycles since softmax start over {B*4} CTAs")
prev = 0.0
labels = {0: "start", 1: "q-tile 0", 2: "q-tile 1", 3: "q-tile 2", 4: "q-tile 3", 20: "finish(3) done"}
for i in [0, 1, 2, 3, 4, 20]:
    v = rel[:, i].mean().item()
    print(f"  {labels[i]:20s} {v:9.0f}  (+{v - prev:7.0f})")
    prev = v
