#!/usr/bin/env python
"""Token kernel times at the bench size (2048 sequences, CUDA events): EMBED, MID (block 1), FINAL through the stage-wise
C-ABI entries.  With T2S_B200_LIB pointing at a knock-out variant (tools/build_variant.py -DT2S_KO_...) this shows what a
resource costs the kernel; the results of such a variant are wrong by construction."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, Workspace, make_dit, stream
from t2ms_b200 import _lib

lib = _lib.load()
model, _ = make_dit(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.randn(B // 2, 64, 30, device=DEV)
emb = torch.randn(B // 2, 128, device=DEV)
t100 = torch.full((1,), 37.0, device=DEV)
out = torch.empty(B, 64, 30, device=DEV)
pk = model.packed()
ws = Workspace(model, B)
lib.t2s_dit_cond(pk.ref, t100.data_ptr(), 0, emb.data_ptr(), 1, B, ws.ptr, stream())
lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream())
lib.t2s_dit_attention(B, ws.ptr, stream())
res = []
for label, fn in (("EMBED", lambda: lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream())),
                  ("MID", lambda: lib.t2s_dit_block_post(pk.ref, 1, B, ws.ptr, stream())),
                  ("FINAL", lambda: lib.t2s_dit_final(pk.ref, out.data_ptr(), B, ws.ptr, stream()))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res.append("%s %.4f" % (label, e0.elapsed_time(e1) / 20))
print("token kernels, %d sequences, ms per launch: " % B + "  ".join(res))
