#!/usr/bin/env python
"""Attention kernel time at the bench size (2048 sequences, CUDA events) + error against the exact fp64 softmax."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, Workspace, make_dit, pack_qkv_images, stream
from t2ms_b200 import _lib

lib = _lib.load()
dit, _ = make_dit(0)
n = 40
g = torch.Generator().manual_seed(1)
q, k, v = (torch.randn(n, 4, 480, 32, generator=g) for _ in range(3))
q = q * 1.5
ref = (torch.softmax((q.double() @ k.double().transpose(-1, -2)) / 32 ** 0.5, -1) @ v.double()).permute(0, 2, 1, 3).reshape(n, 480, 128)
ws = Workspace(dit, n)
ws._view(1, n * 4 * 47104 * 2, torch.float16, (n, 4, 47104)).copy_(pack_qkv_images(q, k, v).to(DEV))
_lib.check(lib.t2s_dit_attention(n, ws.ptr, stream()))
torch.cuda.synchronize()
o = ws.o().double().cpu()
print("rel-L2 vs exact softmax: %.3e" % ((o - ref).norm() / ref.norm()).item())
N = 2048
wsb = Workspace(dit, N)
wsb._view(1, N * 4 * 47104 * 2, torch.float16, (N * 4 * 47104,)).normal_()
for _ in range(3):
    lib.t2s_dit_attention(N, wsb.ptr, stream())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    lib.t2s_dit_attention(N, wsb.ptr, stream())
e1.record()
torch.cuda.synchronize()
print("attention, 2048 sequences: %.4f ms per launch" % (e0.elapsed_time(e1) / 20))
