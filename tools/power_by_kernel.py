#!/usr/bin/env python
"""SM clock and board power while ONE kernel of the sampling step runs back to back for a few seconds (nvidia-smi sampled every
100 ms), and while the whole guided loop runs: which phase of a step the 1000 W power cap throttles.
    python tools/power_by_kernel.py [seconds per leg]"""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import DEV, Workspace, make_dit, stream
from t2ms_b200 import T2SSampler, _lib, synth

SECS = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
lib = _lib.load()
model, _ = make_dit(0)
batch = 1024
B = 2 * batch
x = torch.randn(batch, 64, 30, device=DEV)
emb = torch.randn(batch, 128, device=DEV)
t100 = torch.full((1,), 37.0, device=DEV)
out = torch.empty(B, 64, 30, device=DEV)
pk = model.packed()
ws = Workspace(model, B)
legs = [("attention", lambda: lib.t2s_dit_attention(B, ws.ptr, stream())),
        ("token MID", lambda: lib.t2s_dit_block_post(pk.ref, 1, B, ws.ptr, stream())),
        ("token FINAL", lambda: lib.t2s_dit_final(pk.ref, out.data_ptr(), B, ws.ptr, stream())),
        ("token EMBED", lambda: lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, B, ws.ptr, stream()))]
lib.t2s_dit_cond(pk.ref, t100.data_ptr(), 0, emb.data_ptr(), 1, B, ws.ptr, stream())
for _, fn in legs:
    fn()
torch.cuda.synchronize()


def sample(stop, rows):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        try:
            c, p = r.stdout.strip().split(",")
            rows.append((float(c), float(p)))
        except ValueError:
            pass
        time.sleep(0.1)


def leg(label, fn, unit_ms=None):
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    th.start()
    e0.record()
    while time.time() - t0 < SECS:
        for _ in range(50):
            fn()
        n += 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set(); th.join()
    rows = rows[len(rows) // 3:]                       # the last two thirds: settled
    clk, pw = statistics.median(r[0] for r in rows), statistics.median(r[1] for r in rows)
    print(f"{label:22s} {e0.elapsed_time(e1) / n:8.4f} ms per launch   SM clock {clk:6.0f} MHz   power {pw:6.0f} W   ({len(rows)} samples)", flush=True)


only = os.environ.get("POWER_LEGS", "")
for label, fn in legs:
    if not only or label in only.split(","):
        leg((os.environ.get("POWER_TAG", "") + " " + label).strip(), fn)
if not only:
    smp = T2SSampler(model)
    e = synth.make_text_embeddings(batch, seed=7).to(DEV)
    x0 = torch.randn(batch, 64, 30, device=DEV)
    leg("guided loop", lambda: smp.sample_latent(e, steps=4, noise=x0))
