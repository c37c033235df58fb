"""Kernel time against the number of full waves: nseq = 74 k sequences = k rounds of 148 two-tile token items / 296 attention
CTAs.  T(k) = a + b k separates the fixed cost of a launch (ramp-up, tail) from the steady-state cost of a round."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, Workspace, make_dit, stream  # noqa: E402
from t2ms_b200 import _lib  # noqa: E402

lib = _lib.load()
dit, _ = make_dit(0)
pk = dit.packed()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
res = {}
for k in (1, 2, 4, 8, 16, 27, 28):
    n = 74 * k
    ws = Workspace(dit, n)
    x = torch.randn(n // 2, 64, 30, device=DEV)
    emb = torch.randn(n // 2, 128, device=DEV)
    t100 = torch.full((1,), 37.0, device=DEV)
    out = torch.empty(n, 64, 30, device=DEV)
    calls = {
        "embed": lambda: lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, n, ws.ptr, stream()),
        "attn": lambda: lib.t2s_dit_attention(n, ws.ptr, stream()),
        "mid": lambda: lib.t2s_dit_block_post(pk.ref, 1, n, ws.ptr, stream()),
        "final": lambda: lib.t2s_dit_final(pk.ref, out.data_ptr(), n, ws.ptr, stream()),
    }
    lib.t2s_dit_cond(pk.ref, t100.data_ptr(), 0, emb.data_ptr(), 1, n, ws.ptr, stream())
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[(name, k)] = sorted(ts)[len(ts) // 2] * 1e3
for name in ("embed", "attn", "mid", "final"):
    print(name, " ".join(f"k={k}: {res[(name, k)]:.1f} us" for k in (1, 2, 4, 8, 16, 27, 28)),
          f"| per round (k 8->16): {(res[(name, 16)] - res[(name, 8)]) / 8:.2f} us, fixed (k=1 minus one round): {res[(name, 1)] - (res[(name, 16)] - res[(name, 8)]) / 8:.1f} us")
