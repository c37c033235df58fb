"""A/B of the fused per-step kernel (csrc/dit_fused.cuh) against the per-phase kernels on the GPU: parity of one forward and
of a short guided loop at several batch sizes, per-step time at the bench workload, scheduler statistics.
    python tools/fused_check.py [--quick] [--inflight N]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2ms_b200 import T2SSampler, Transformer, _lib, synth, vqvae  # noqa: E402
from t2ms_b200.compat import VAE_ARGS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--inflight", type=int, default=0)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = "cuda:0"
lib = _lib.load()
dit = Transformer(); dit.load_state_dict(synth.make_dit_state(0, bias_std=0.02)); dit = dit.to(dev).eval()
vae = vqvae(VAE_ARGS); vae.load_state_dict(synth.make_vae_state(1)); vae = vae.to(dev).eval()
smp = T2SSampler(dit, vae)
res = {}


def fwd(B, fused):
    lib.t2s_set_fused(1 if fused else -1, a.inflight)
    x, emb = synth.make_noise(B, seed=3).to(dev), synth.make_text_embeddings(B, seed=4).to(dev)
    t = torch.linspace(0, 0.9, B, device=dev)
    with torch.no_grad():
        o = dit(input=x, t=t, text_input=emb)
    torch.cuda.synchronize()
    return o


for B in ([3, 80] if a.quick else [1, 3, 16, 80, 513]):
    o0, o1 = fwd(B, False), fwd(B, True)
    d = (o0 - o1).abs().max().item()
    print(f"forward B={B}: fused vs per-phase max-abs diff {d:.3e} (|out| max {o0.abs().max().item():.3f}) finite={torch.isfinite(o1).all().item()}", flush=True)
    res[f"fwd_diff_B{B}"] = d

for B in ([5, 64] if a.quick else [2, 5, 64, 300]):
    emb, x0 = synth.make_text_embeddings(B, seed=5).to(dev), synth.make_noise(B, seed=6).to(dev)
    outs = []
    for fused in (False, True):
        lib.t2s_set_fused(1 if fused else -1, a.inflight)
        outs.append(smp.sample(emb, 96, steps=4, noise=x0))
        torch.cuda.synchronize()
    d = (outs[0] - outs[1]).abs().max().item()
    print(f"RF 4 steps B={B}: fused vs per-phase series max-abs diff {d:.3e}", flush=True)
    res[f"rf_diff_B{B}"] = d
    lib.t2s_set_fused(1, a.inflight)
    again = smp.sample(emb, 96, steps=4, noise=x0)
    print(f"   fused run-to-run identical: {torch.equal(again, outs[1])}", flush=True)

# timing at the bench workload
B = a.batch
emb, x0 = synth.make_text_embeddings(B, seed=7).to(dev), torch.randn(B, 64, 30, device=dev)
stats = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
for name, mp in (("per_phase", -1), ("fused", 1)):
    lib.t2s_set_fused(mp, a.inflight)
    for _ in range(2):
        smp.sample(emb, 96, steps=a.steps, noise=x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        smp.sample(emb, 96, steps=a.steps, noise=x0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3 / a.steps
    print(f"{name}: {ms:.3f} ms per guided step at batch {B} -> {B / (ms * 100 / 1e3):.0f} series/s at 100 steps", flush=True)
    res[f"ms_per_step_{name}"] = ms
lib.t2s_set_fused(1, a.inflight)
lib.t2s_debug_set_fused_stats(stats.data_ptr())
smp.sample(emb, 96, steps=1, noise=x0)
torch.cuda.synchronize()
lib.t2s_debug_set_fused_stats(None)
s = stats.view(148, 8).cpu()
tot = s[:, 4].float()
print("fused stats (one step): token items/CTA min %d max %d; attention units/CTA min %d max %d" % (s[:, 0].min(), s[:, 0].max(), s[:, 2].min(), s[:, 2].max()))
print("   starved share of the kernel: token %.1f %%, attention %.1f %%; kernel cycles %.0f" % (100 * (s[:, 1].float() / tot).mean(), 100 * (s[:, 3].float() / tot).mean(), tot.mean()))
res["token_starved"] = (s[:, 1].float() / tot).mean().item()
res["attn_starved"] = (s[:, 3].float() / tot).mean().item()
res["kernel_cycles"] = tot.mean().item()
# phase trace of the token epilogue (row 0, half 0): MID items of every CTA, items 4..15 (steady state)
trace = torch.zeros(148 * 16 * 32, dtype=torch.int64, device=dev)
lib.t2s_set_fused(1, a.inflight)
lib.t2s_debug_set_fused_trace(trace.data_ptr())
smp.sample(emb, 96, steps=1, noise=x0)
torch.cuda.synchronize()
lib.t2s_debug_set_fused_trace(None)
t = trace.view(148, 16, 32).cpu()
names = ["desc", "vec", "hq/acc0 wait", "resid+ln1 -> A2", "acc1 wait", "gelu a", "acc2 wait", "gelu b", "acc3 wait", "resid2+merge", "ln' -> A3",
         "acc4 wait", "q store", "acc5 wait", "k store", "acc6 wait", "v store", "end barrier", "tok_done"]
for mode, label in ((1, "MID"), (0, "EMBED"), (2, "FINAL")):
    sel = [(c, i) for c in range(148) for i in range(3, 15) if t[c, i, 31] // 16 == mode and t[c, i, 31] > 0 and t[c, i + 1, 0] > 0]
    if not sel:
        continue
    d = torch.stack([t[c, i, 1:20] - t[c, i, 0:19] for c, i in sel]).float()
    d = torch.where((torch.stack([t[c, i, 1:20] for c, i in sel]) > 0) & (torch.stack([t[c, i, 0:19] for c, i in sel]) > 0), d, torch.zeros_like(d))
    tot = torch.tensor([float(t[c, i + 1, 0] - t[c, i, 0]) for c, i in sel])
    print(f"{label}: {len(sel)} items, {tot.mean():.0f} cycles per item (median {tot.median():.0f})")
    print("   " + ", ".join(f"{n} {v:.0f}" for n, v in zip(names, d.mean(0).tolist()) if v > 0))
    res[f"trace_{label}"] = {"cycles": tot.mean().item(), **{n: v for n, v in zip(names, d.mean(0).tolist())}}
lib.t2s_set_fused(-1, 0)
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
