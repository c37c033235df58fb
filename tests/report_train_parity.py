#!/usr/bin/env python
"""Training-step report on one B200: per-tensor gradient error against the fp32 CPU oracle (train.py:83-85) and
the device time of the fused step (forward + MSE + backward + AdamW) at several batch sizes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import t2s_oracle as O
from t2ms_b200 import Transformer, synth
from t2ms_b200.training import DitTrainer, trainable_names

DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    out = {}
    dsd = synth.make_dit_state(15, bias_std=0.02)
    B = 4
    x1, x0, emb = synth.make_noise(B, seed=5), synth.make_noise(B, seed=6), synth.make_text_embeddings(B, seed=7)
    t = torch.tensor([0.1, 0.5, 0.73, 1.0])
    x_t, target = O.rf_create_flow(x1, t, x0), x1 - x0
    loss_ref, grads_ref = O.train_step_grads(dsd, x_t, t, emb, target)
    m = Transformer()
    m.load_state_dict(dsd)
    m = m.to(DEV).train()
    tr = DitTrainer(m)
    tr.zero_grad()
    pred = torch.empty(B, 64, 30, device=DEV)
    tr.forward_backward(x_t.to(DEV), t.to(DEV), emb.to(DEV), target.to(DEV), pred=pred)
    torch.cuda.synchronize()
    errs = {n: rel(tr.grads.view(n), grads_ref[n]) for n in trainable_names()}
    out["parity"] = {"pred_rel_l2": rel(pred, O.dit_forward(dsd, x_t, t, emb)),
                     "loss_rel": abs(tr.loss_sum.item() / (B * 1920) - loss_ref.item()) / loss_ref.item(),
                     "grad_rel_l2_max": max(errs.values()), "grad_rel_l2_median": sorted(errs.values())[len(errs) // 2],
                     "worst": max(errs, key=errs.get)}
    out["timing"] = []
    for B in (64, 256, 1024):
        x_t = torch.randn(B, 64, 30, device=DEV)
        tgt = torch.randn(B, 64, 30, device=DEV)
        tt = torch.rand(B, device=DEV)
        e = torch.nn.functional.normalize(torch.randn(B, 128, device=DEV), dim=-1)
        for _ in range(2):
            tr.step(x_t, tt, e, tgt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            tr.step(x_t, tt, e, tgt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flop = 3 * 976_960_512 * B
        out["timing"].append({"batch": B, "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1),
                              "tflops_algorithmic(3x fwd)": round(flop / ms / 1e9, 1)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
