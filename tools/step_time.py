"""ms per guided step at a batch size, A/B over the launch switches: python tools/step_time.py [--batch 1024] [--steps 20]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2ms_b200 import T2SSampler, Transformer, _lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = "cuda:0"
lib = _lib.load()
dit = Transformer(); dit.load_state_dict(synth.make_dit_state(0)); dit = dit.to(dev).eval()
smp = T2SSampler(dit)
emb, x0 = synth.make_text_embeddings(a.batch, seed=7).to(dev), torch.randn(a.batch, 64, 30, device=dev)


def run(label):
    for _ in range(2):
        smp.sample_latent(emb, steps=a.steps, noise=x0)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        smp.sample_latent(emb, steps=a.steps, noise=x0)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    print(f"{label}: {best:.4f} ms per guided step at batch {a.batch} -> {a.batch / (best * 100 / 1e3):.0f} series/s at 100 steps", flush=True)
    return best


for rnd in range(2):
    lib.t2s_set_pdl(0); run("plain launches  ")
    lib.t2s_set_pdl(1); run("PDL launches    ")
