#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference is mounted:

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference imports two packages that are absent from this image:
``timm==1.0.11`` (model/denoiser/transformer.py:3) and ``matplotlib`` (rectified_flow.py:3,
DDPM.py:3).  Two throw-away shims are written to a temp dir and put on sys.path: an empty
``matplotlib.pyplot`` and a ``timm.models.vision_transformer`` that restates the published
timm 1.0.11 ``Attention`` / ``Mlp`` forward (qkv Linear -> (B,N,3,H,d) -> SDPA -> proj;
fc1 -> act -> fc2).  Nothing from /root/reference is copied into the repo; only the numeric
outputs are committed, together with this script.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import textwrap
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200 import synth  # noqa: E402

TIMM_SHIM = '''
import torch, torch.nn as nn, torch.nn.functional as F
class PatchEmbed(nn.Module):
    pass
class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, **kw):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity(); self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)
    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0., **kw):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)
    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))
'''


def install_shims(ref_root: str):
    d = tempfile.mkdtemp(prefix="t2s_shims_")
    os.makedirs(os.path.join(d, "timm", "models"))
    os.makedirs(os.path.join(d, "matplotlib"))
    open(os.path.join(d, "timm", "__init__.py"), "w").close()
    open(os.path.join(d, "timm", "models", "__init__.py"), "w").close()
    with open(os.path.join(d, "timm", "models", "vision_transformer.py"), "w") as f:
        f.write(textwrap.dedent(TIMM_SHIM))
    open(os.path.join(d, "matplotlib", "__init__.py"), "w").close()
    open(os.path.join(d, "matplotlib", "pyplot.py"), "w").close()
    open(os.path.join(d, "matplotlib", "animation.py"), "w").close()
    sys.path.insert(0, d)
    sys.path.insert(0, ref_root)


def load_reference(ref_root: str):
    install_shims(ref_root)
    from model.denoiser.transformer import Transformer, TimeEmbedding
    from model.backbone.rectified_flow import RectifiedFlow
    from model.backbone.DDPM import DDPM
    from model.pretrained.vqvae import vqvae
    return Transformer, TimeEmbedding, RectifiedFlow, DDPM, vqvae


VAE_ARGS = Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256, embedding_dim=64)


def build_models(ref, dit_seed, vae_seed, bias_std):
    Transformer, _, _, _, vqvae = ref
    dit_sd = synth.make_dit_state(dit_seed, bias_std=bias_std)
    vae_sd = synth.make_vae_state(vae_seed)
    dit = Transformer()
    dit.load_state_dict(dit_sd, strict=True)
    dit.eval()
    vae = vqvae(VAE_ARGS)
    vae.load_state_dict(vae_sd, strict=True)
    vae.eval()
    return dit, vae, dit_sd, vae_sd


def npf(t):
    return t.detach().cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    torch.set_num_threads(8)
    ref = load_reference(a.ref)
    Transformer, TimeEmbedding, RectifiedFlow, DDPM, vqvae = ref

    # ---------------------------------------------------------------- DiT forward
    dit_seed, vae_seed, bias_std = 11, 12, 0.05
    dit, vae, dit_sd, vae_sd = build_models(ref, dit_seed, vae_seed, bias_std)
    B = 3
    x = synth.make_noise(B, seed=21)
    emb = synth.make_text_embeddings(B, seed=22)
    t_f = torch.tensor([0.0, 0.37, 0.99], dtype=torch.float32)
    t_i = torch.tensor([0, 417, 999], dtype=torch.long)
    out = {}
    with torch.no_grad():
        out["temb_float"] = npf(TimeEmbedding(128)(t_f))
        out["temb_int"] = npf(TimeEmbedding(128)(t_i))
        out["cond_float"] = npf(dit(input=x, t=t_f, text_input=emb))
        out["uncond_float"] = npf(dit(input=x, t=t_f, text_input=None))
        out["cond_int"] = npf(dit(input=x, t=t_i, text_input=emb))
        # hidden states after each block (token slices keep the fixture small)
        hs = []
        hooks = [l.register_forward_hook(lambda m, i, o: hs.append(o)) for l in dit.layers]
        dit(input=x, t=t_f, text_input=emb)
        for h in hooks:
            h.remove()
        tok = [0, 1, 31, 32, 239, 240, 478, 479]
        out["hidden_tok"] = np.array(tok)
        out["hidden"] = np.stack([npf(h[:, tok, :]) for h in hs])          # (4,B,8,128)
    np.savez_compressed(
        os.path.join(a.out, "dit_forward.npz"), dit_seed=dit_seed, bias_std=bias_std,
        dit_checksum=synth.state_checksum(dit_sd), x=npf(x), emb=npf(emb), t_float=npf(t_f), t_int=npf(t_i), **out)

    # reference-default init (adaLN zero => identity blocks, cond == uncond bit-exactly)
    torch.manual_seed(5)
    dit0 = Transformer().eval()
    with torch.no_grad():
        o0c = dit0(input=x, t=t_f, text_input=emb)
        o0u = dit0(input=x, t=t_f, text_input=None)
    assert torch.equal(o0c, o0u)

    # ---------------------------------------------------------------- VAE
    out = {}
    with torch.no_grad():
        for L in (24, 48, 96):
            s = synth.make_series(3, L, seed=30 + L)
            z, before = vae.encoder(s)
            rec, after = vae.decoder(z, length=L)
            out[f"series_{L}"] = npf(s)
            out[f"z_{L}"] = npf(z)
            out[f"before_{L}"] = npf(before)
            out[f"rec_{L}"] = npf(rec)
            out[f"after_{L}"] = npf(after)
        zlat = synth.make_noise(2, seed=33)
        for L in (24, 48, 96):
            rec, after = vae.decoder(zlat, length=L)
            out[f"dec_noise_{L}"] = npf(rec)
        rec1, _ = vae.decoder(zlat[:1], length=48)                  # torch.squeeze drops the batch dim
        out["dec_b1_48"] = npf(rec1)
    np.savez_compressed(os.path.join(a.out, "vae.npz"), vae_seed=vae_seed,
                        vae_checksum=synth.state_checksum(vae_sd), zlat=npf(zlat), **out)

    # ---------------------------------------------------------------- backbone maths
    rf = RectifiedFlow()
    ddpm = DDPM(1000, "cpu")
    x1 = synth.make_noise(4, seed=41) * 0.5
    tt = torch.tensor([0.0, 0.25, 0.5, 1.0])
    torch.manual_seed(77)
    st = torch.get_rng_state()
    x0_expected = torch.randn_like(x1)
    torch.set_rng_state(st)
    x_t, x_0 = rf.create_flow(x1, tt)
    assert torch.equal(x_0, x0_expected)
    ti = torch.tensor([0, 10, 500, 999])
    eps = synth.make_noise(4, seed=42)
    q, _ = ddpm.q_sample(x1, ti, eps)
    st = torch.get_rng_state()
    pn = torch.randn(x1.shape)
    torch.set_rng_state(st)
    p = ddpm.p_sample(x1, eps, ti)
    np.savez_compressed(
        os.path.join(a.out, "backbone.npz"), x1=npf(x1), t=npf(tt), x0=npf(x_0), x_t=npf(x_t),
        euler=npf(rf.euler(x1, eps, 0.01)), rf_loss=npf(rf.loss(x1, eps)), ti=npf(ti), eps=npf(eps),
        q_sample=npf(q), p_noise=npf(pn), p_sample=npf(p), beta=npf(ddpm.beta), alpha=npf(ddpm.alpha),
        alpha_bar=npf(ddpm.alpha_bar), ddpm_loss=npf(ddpm.loss(x1, eps)))

    # ---------------------------------------------------------------- sampling loops (infer.py:75-95)
    dit, vae, dit_sd, vae_sd = build_models(ref, 13, 14, 0.02)
    out = {}
    Bs = 2
    emb = synth.make_text_embeddings(Bs, seed=52)
    with torch.no_grad():
        for steps, L, cfg in ((4, 24, 7.0), (6, 48, 5.0), (5, 96, 7.0)):
            noise = synth.make_noise(Bs, seed=50 + steps)
            x_t = noise.clone()
            vels = []
            for j in range(steps):
                t = torch.round(torch.full((x_t.shape[0],), j * 1.0 / steps) * steps) / steps
                pu = dit(input=x_t, t=t, text_input=None)
                pc = dit(input=x_t, t=t, text_input=emb)
                pred = pu + cfg * (pc - pu)
                vels.append(npf(pred))
                x_t = rf.euler(x_t, pred, 1.0 / steps)
            ser, _ = vae.decoder(x_t, length=L)
            out[f"rf_{steps}_{L}_noise"] = npf(noise)
            out[f"rf_{steps}_{L}_vel"] = np.stack(vels)
            out[f"rf_{steps}_{L}_latent"] = npf(x_t)
            out[f"rf_{steps}_{L}_series"] = npf(ser)
            out[f"rf_{steps}_{L}_cfg"] = cfg
        # DDPM, infer.py:83-88
        steps, L, cfg = 8, 48, 7.0
        ddpm = DDPM(steps, "cpu")
        noise = synth.make_noise(Bs, seed=60)
        x_t = noise.clone()
        torch.manual_seed(99)
        step_noise, eps_l = [], []
        import math
        for j in range(steps):
            t = torch.full((x_t.size(0),), math.floor(steps - 1 - j), dtype=torch.long)
            pu = dit(input=x_t, t=t, text_input=None)
            pc = dit(input=x_t, t=t, text_input=emb)
            pred = pu + cfg * (pc - pu)
            eps_l.append(npf(pred))
            st = torch.get_rng_state()
            step_noise.append(npf(torch.randn(x_t.shape)))
            torch.set_rng_state(st)
            x_t = ddpm.p_sample(x_t, pred, t)
        ser, _ = vae.decoder(x_t, length=L)
        out["ddpm_noise"] = npf(noise)
        out["ddpm_step_noise"] = np.stack(step_noise)
        out["ddpm_eps"] = np.stack(eps_l)
        out["ddpm_latent"] = npf(x_t)
        out["ddpm_series"] = npf(ser)
    np.savez_compressed(
        os.path.join(a.out, "sampling.npz"), dit_seed=13, vae_seed=14, bias_std=0.02,
        dit_checksum=synth.state_checksum(dit_sd), vae_checksum=synth.state_checksum(vae_sd),
        emb=npf(emb), ddpm_steps=8, ddpm_L=48, ddpm_cfg=7.0, **out)

    # ---------------------------------------------------------------- training step (train.py:66-87)
    torch.manual_seed(123)
    dit, vae, dit_sd, vae_sd = build_models(ref, 15, 16, 0.02)
    dit.train()
    dit.encoder = vae.encoder
    for n, p_ in dit.named_parameters():
        if "encoder" in n:
            p_.requires_grad = False
    opt = torch.optim.AdamW(dit.parameters(), lr=1e-4, weight_decay=0.0)
    Bt, L = 4, 48
    series = synth.make_series(Bt, L, seed=70)
    emb = synth.make_text_embeddings(Bt, seed=71)
    x0 = synth.make_noise(Bt, seed=72)
    t = torch.tensor([0.1, 0.5, 0.73, 1.0])
    with torch.no_grad():
        x1, _ = dit.encoder(series)
    x_t = t[:, None, None] * x1 + (1 - t[:, None, None]) * x0
    target = x1 - x0
    opt.zero_grad()
    pred = dit(input=x_t, t=t, text_input=emb)
    loss = rf.loss(pred, target)
    loss.backward()
    gsel = {}
    for n, p_ in dit.named_parameters():
        if p_.grad is not None:
            gsel["grad_norm/" + n] = float(p_.grad.norm())
    probe = ["layers.0.attn.qkv.weight", "layers.3.mlp.fc2.weight", "layers.1.adaLN_modulation.1.bias",
             "patch_emb.weight", "conv.weight", "linear_emb_to_patch.weight", "ln.weight"]
    gfull = {"grad/" + n: npf(dict(dit.named_parameters())[n].grad).reshape(-1)[:256] for n in probe}
    opt.step()
    pfull = {"param_after/" + n: npf(dict(dit.named_parameters())[n]).reshape(-1)[:256] for n in probe}
    np.savez_compressed(
        os.path.join(a.out, "train.npz"), dit_seed=15, vae_seed=16, bias_std=0.02,
        dit_checksum=synth.state_checksum(dit_sd), series=npf(series), emb=npf(emb), x0=npf(x0), t=npf(t),
        x1=npf(x1), loss=float(loss), names=np.array(list(gsel.keys())), norms=np.array(list(gsel.values())),
        **gfull, **pfull)

    for f in sorted(os.listdir(a.out)):
        print(f, os.path.getsize(os.path.join(a.out, f)))


if __name__ == "__main__":
    main()
