#!/usr/bin/env python
"""Golden fixtures for the rows SURVEY.md §8(f) marks "next", generated from the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference is mounted):

    python oracle/make_golden_next.py [--only eval,dit_tokens,vae_train]

Writes tests/golden/eval.npz (evaluation.py calculate_mse / calculate_wape; the module cannot be imported here — it pulls
dtaidistance and TS2Vec at import time — so the two function definitions are compiled from the reference file's own
source text, unmodified, via ``ast``), and further fixtures as the corresponding rows are built.  Nothing from
/root/reference is copied into the repo; only numeric outputs are committed, together with this script.
"""
from __future__ import annotations

import argparse
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def reference_functions(path: str, names):
    """Compile the named top-level functions of a reference file from its own source (no module import)."""
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), [n.name for n in keep]
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def make_eval(ref: str, out: str):
    mse, wape = reference_functions(os.path.join(ref, "evaluation.py"), ["calculate_mse", "calculate_wape"])
    rng = np.random.default_rng(5)
    cases = {}
    for name, (n, L) in {"a": (7, 24), "b": (33, 96), "c": (5, 48)}.items():
        ori = rng.random((n, L, 1)).astype(np.float32)                 # the saved layout (infer.py:115-116)
        gen = (ori + 0.1 * rng.standard_normal((n, L, 1))).astype(np.float32)
        if name == "c":
            ori[2] = 0.0                                               # a sample with a zero denominator -> NaN -> nanmean skips it
        o, g = np.transpose(ori, (0, 2, 1)), np.transpose(gen, (0, 2, 1))   # evaluation.py:295-296
        cases[f"{name}/ori"], cases[f"{name}/gen"] = ori, gen
        cases[f"{name}/mse"], cases[f"{name}/wape"] = np.float64(mse(o, g)), np.float64(wape(o, g))
    np.savez_compressed(os.path.join(out, "eval.npz"), **cases)


def make_dit_tokens(ref: str, out: str):
    """The fork's variable-width denoiser, model/denoiser/mytransformer.py ``Transformer(dim)`` with dim = 50 and 64
    (config.yaml:46,91 flow_dim): forward (float and int64 t, with / without text) and a 3-step guided rectified-flow and
    DDPM loop (infer.py:75-88 semantics, as myinfer.py drives it) on the same synthetic weights / inputs as the tests."""
    from make_golden import install_shims
    from t2ms_b200 import synth
    install_shims(ref)
    sys.path.insert(0, ref)
    from model.denoiser.mytransformer import Transformer
    from model.backbone.rectified_flow import RectifiedFlow
    from model.backbone.DDPM import DDPM
    res = {}
    for dim in (50, 64):
        sd = synth.make_dit_state(40 + dim, bias_std=0.05, dim=dim)
        m = Transformer(dim)
        m.load_state_dict(sd, strict=True)
        m.eval()
        B = 2
        x = synth.make_noise(B, seed=50 + dim, dim=dim)
        emb = synth.make_text_embeddings(B, seed=60 + dim)
        t_f, t_i = torch.tensor([0.25, 0.9]), torch.tensor([3, 871], dtype=torch.long)
        k = f"h{dim}/"
        with torch.no_grad():
            res[k + "cond_float"] = m(input=x, t=t_f, text_input=emb).numpy()
            res[k + "uncond_float"] = m(input=x, t=t_f, text_input=None).numpy()
            res[k + "cond_int"] = m(input=x, t=t_i, text_input=emb).numpy()
            steps, cfg = 3, 7.0
            rf, x_t, vel = RectifiedFlow(), x.clone(), []
            for j in range(steps):
                t = torch.round(torch.full((B,), j * 1.0 / steps) * steps) / steps
                u, c = m(input=x_t, t=t, text_input=None), m(input=x_t, t=t, text_input=emb)
                pred = u + cfg * (c - u)
                vel.append(pred.numpy())
                x_t = rf.euler(x_t, pred, 1.0 / steps)
            res[k + "rf_vel"], res[k + "rf_final"] = np.stack(vel), x_t.numpy()
            ddpm, x_t, eps = DDPM(steps, "cpu"), x.clone(), []
            sn = synth.make_step_noise(steps, B, seed=70 + dim, dim=dim)
            for j in range(steps):
                t = torch.full((B,), steps - 1 - j, dtype=torch.long)
                u, c = m(input=x_t, t=t, text_input=None), m(input=x_t, t=t, text_input=emb)
                pred = u + cfg * (c - u)
                eps.append(pred.numpy())
                torch.manual_seed(0)
                real_randn = torch.randn
                torch.randn = lambda *a, **kw: sn[j].clone()            # DDPM.p_sample draws its noise inside (DDPM.py:35)
                try:
                    x_t = ddpm.p_sample(x_t, pred, t)
                finally:
                    torch.randn = real_randn
            res[k + "ddpm_eps"], res[k + "ddpm_final"] = np.stack(eps), x_t.numpy()
        res[k + "checksum"] = np.array(synth.state_checksum(sd))
    np.savez_compressed(os.path.join(out, "dit_tokens.npz"), **res)


def make_dit_tokens_train(ref: str, out: str):
    """One rectified-flow training step of the fork's variable-width denoiser (mytrain.py:66-87 around
    model/denoiser/mytransformer.py Transformer(dim), dim = 50 and 64): loss and every parameter gradient (norms + slices)
    on synthetic latents (the multivariate LA-VAE encoder is exercised separately)."""
    from make_golden import install_shims
    from t2ms_b200 import synth
    install_shims(ref)
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from model.denoiser.mytransformer import Transformer
    from model.backbone.rectified_flow import RectifiedFlow
    res = {}
    for dim in (50, 64):
        sd = synth.make_dit_state(140 + dim, bias_std=0.02, dim=dim)
        m = Transformer(dim)
        m.load_state_dict(sd, strict=True)
        m.train()
        B = 3
        x1, x0 = synth.make_noise(B, seed=150 + dim, dim=dim), synth.make_noise(B, seed=160 + dim, dim=dim)
        emb = synth.make_text_embeddings(B, seed=170 + dim)
        t = torch.tensor([0.2, 0.55, 0.9])
        rf = RectifiedFlow()
        x_t = t[:, None, None] * x1 + (1 - t[:, None, None]) * x0           # create_flow with the noise fixed (rectified_flow.py:8-12)
        target = x1 - x0
        pred = m(input=x_t, t=t, text_input=emb)
        loss = rf.loss(pred, target)
        loss.backward()
        k = f"h{dim}/"
        res[k + "loss"] = np.float64(loss.item())
        res[k + "pred"] = pred.detach().numpy()
        names = [n for n, p in m.named_parameters() if p.grad is not None]
        res[k + "names"] = np.array(names)
        res[k + "grad_norms"] = np.array([float(dict(m.named_parameters())[n].grad.norm()) for n in names])
        for n in names:
            res[k + "grad/" + n] = dict(m.named_parameters())[n].grad.reshape(-1)[:128].numpy().copy()
        res[k + "checksum"] = np.array(synth.state_checksum(sd))
    np.savez_compressed(os.path.join(out, "dit_tokens_train.npz"), **res)


def make_vae_train(ref: str, out: str):
    """LA-VAE training step: vqvae.shared_eval(batch, optimizer, 'train') of model/pretrained/vqvae.py:118-135 (univariate,
    L = 48) and of the fork's model/pretrained/myvqvae.py:116-136 (input_dim 7, flow_dim 50, L = 100 and L = 90 — the latter
    exercises the final interpolation of myvqvae.py:85), with the reference's own optimizer (core.py:15-20 AdamW lr 1e-3,
    weight_decay 1e-2).  Saved: losses, recon / z, every parameter's gradient norm, gradient slices, parameters after the step."""
    from argparse import Namespace
    from make_golden import install_shims
    from t2ms_b200 import synth
    install_shims(ref)
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from model.pretrained.vqvae import vqvae as vq_uni
    from model.pretrained.myvqvae import vqvae as vq_multi
    res = {}
    cases = {"uni48": (vq_uni, dict(), 1, 30, 48, 4), "multi100": (vq_multi, dict(flow_dim=50, input_dim=7), 7, 50, 100, 3),
             "multi90": (vq_multi, dict(flow_dim=50, input_dim=7), 7, 50, 90, 2)}
    for name, (cls, extra, cin, flow, L, B) in cases.items():
        args = Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256, embedding_dim=64, **extra)
        m = cls(args)
        sd = synth.make_vae_state(80 + L, in_channels=cin)
        m.load_state_dict(sd, strict=True)
        m.train()
        g = torch.Generator().manual_seed(90 + L)
        batch = torch.rand(B, L, generator=g) if cin == 1 else torch.rand(B, cin, L, generator=g)
        opt, _ = m.configure_optimizers(lr=1e-3)
        loss, recon_error, recon, z = m.shared_eval(batch, opt, "train")
        k = name + "/"
        res[k + "batch"], res[k + "loss"], res[k + "recon_error"] = batch.numpy(), np.float64(loss.item()), np.float64(recon_error.item())
        res[k + "recon"], res[k + "z"] = recon.detach().numpy(), z.detach().numpy()
        names = [n for n, _ in m.named_parameters()]
        res[k + "names"] = np.array(names)
        res[k + "grad_norms"] = np.array([float(p.grad.norm()) for _, p in m.named_parameters()])
        for n, p in m.named_parameters():
            res[k + "grad/" + n] = p.grad.reshape(-1)[:128].numpy().copy()
            res[k + "after/" + n] = p.detach().reshape(-1)[:128].numpy().copy()
        res[k + "checksum"] = np.array(synth.state_checksum(sd))
        with torch.no_grad():
            m.eval()
            l2, r2, _, _ = m.shared_eval(batch, opt, "val")
            res[k + "val_loss_after"] = np.float64(l2.item())
    np.savez_compressed(os.path.join(out, "vae_train.npz"), **res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    only = set(filter(None, a.only.split(",")))
    torch.set_num_threads(8)
    makers = {"eval": make_eval, "dit_tokens": make_dit_tokens, "dit_tokens_train": make_dit_tokens_train, "vae_train": make_vae_train}
    for name, fn in makers.items():
        if not only or name in only:
            fn(a.ref, a.out)
            print("wrote", name)
    for f in sorted(os.listdir(a.out)):
        print(f, os.path.getsize(os.path.join(a.out, f)))


if __name__ == "__main__":
    main()
