"""CPU oracle for the T2S generation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-tensor (torch fp32, CPU) restatement of the reference algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg
may import it; nothing under ``t2ms_b200/`` does (the product path is CUDA only and fails loudly
when the extension is missing).

Parity pinning: the reference (Bill9125/T2MS) ships no tests and no golden vectors, so
this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container:
``oracle/make_golden.py`` imports the unmodified reference modules from /root/reference (with two
import shims for the absent third-party ``timm==1.0.11`` Attention/Mlp and ``matplotlib``) and
writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against
those fixtures.

Third-party arithmetic restated here: ``timm.models.vision_transformer.Attention`` and ``Mlp``
(timm==1.0.11, requirements.txt:9), called from model/denoiser/transformer.py:104-105,116-117.

All ``file:line`` citations are relative to the reference repo root.  Weights are passed as a plain
``dict[str, Tensor]`` using the reference's state-dict key names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------------------------
# DiT denoiser (model/denoiser/transformer.py)
# ----------------------------------------------------------------------------------------------

D_MODEL = 128      # transformer.py:97,135
N_HEADS = 4        # transformer.py:104
N_LAYERS = 4       # transformer.py:149
LAT_C = 64         # latent channels  (self.W, transformer.py:134)
LAT_P = 30         # latent positions (self.H, transformer.py:133)
N_TOK = 480        # (30/2)*(64/2), transformer.py:136


def modulate(x, shift, scale):
    """transformer.py:7-8"""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def sinusoidal_pos_embed(num_positions: int, d_model: int) -> torch.Tensor:
    """transformer.py:14-23 -> (1, num_positions, d_model)"""
    position = torch.arange(num_positions).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model)).unsqueeze(0)
    pe = torch.zeros(num_positions, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def time_embedding(t: torch.Tensor, dim: int = D_MODEL) -> torch.Tensor:
    """TimeEmbedding.forward, transformer.py:30-40.  t: (B,) float32 or int64 -> (B, dim)."""
    t = t * 100.0
    t = t.unsqueeze(-1)
    freqs = torch.pow(10000, torch.linspace(0, 1, dim // 2)).to(t.device)
    sin_emb = torch.sin(t[:, None] / freqs)
    cos_emb = torch.cos(t[:, None] / freqs)
    return torch.cat([sin_emb, cos_emb], dim=-1).squeeze(1)


def attention(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """timm==1.0.11 Attention.forward (qkv_bias=True, no qk-norm, no dropout), call site
    transformer.py:116.  softmax(q k^T / sqrt(d_head)) v, heads re-concatenated, then proj."""
    B, N, C = x.shape
    hd = C // N_HEADS
    qkv = F.linear(x, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, N_HEADS, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    att = (q * hd ** -0.5) @ k.transpose(-2, -1)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])


def mlp(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """timm==1.0.11 Mlp.forward with act=GELU(tanh), drop=0, call site transformer.py:117."""
    x = F.linear(x, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])
    x = F.gelu(x, approximate="tanh")
    return F.linear(x, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def dit_layer(sd: SD, l: int, x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """Transformerlayer.forward, transformer.py:114-124 (norms: :102-103, eps 1e-6, no affine)."""
    pre = f"layers.{l}."
    ada = F.linear(F.silu(c), sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
    sh1, sc1, g1, sh2, sc2, g2 = ada.chunk(6, dim=1)
    x = x + g1.unsqueeze(1) * attention(sd, pre, modulate(F.layer_norm(x, (D_MODEL,), eps=1e-6), sh1, sc1))
    x = x + g2.unsqueeze(1) * mlp(sd, pre, modulate(F.layer_norm(x, (D_MODEL,), eps=1e-6), sh2, sc2))
    return x


def dit_embed(sd: SD, inp: torch.Tensor) -> torch.Tensor:
    """Patchify + embed, transformer.py:166-172.  inp (B,64,H) -> (B,16 H,128); H = 30 in T2S, H = dim in the fork's
    Transformer(dim) (mytransformer.py:128-136,166-172: identical code with self.H = dim)."""
    x = inp.permute(0, 2, 1).unsqueeze(1)                      # (B,1,H,64)
    x = F.conv2d(x, sd["conv.weight"], sd["conv.bias"], stride=2)   # (B,4,H/2,32)
    x = x.permute(0, 2, 3, 1)
    x = x.reshape(x.size(0), x.size(1) * x.size(2), x.size(3))
    x = F.linear(x, sd["patch_emb.weight"], sd["patch_emb.bias"])
    return x + sd["pos_embed"]


def dit_unembed(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """Final LN + projection + unpatchify, transformer.py:182-190 (mytransformer.py:182-190 with H = dim).
    (B,16 H,128) -> (B,64,H)."""
    x = F.layer_norm(x, (D_MODEL,), sd["ln.weight"], sd["ln.bias"], eps=1e-5)
    x = F.linear(x, sd["linear_emb_to_patch.weight"], sd["linear_emb_to_patch.bias"])
    B = x.size(0)
    lat_p = x.size(1) // (LAT_C // 2) * 2                      # H
    x = x.view(B, lat_p // 2, LAT_C // 2, 1, 2, 2)
    x = x.permute(0, 3, 1, 2, 4, 5).permute(0, 1, 2, 4, 3, 5)
    x = x.reshape(B, 1, lat_p, LAT_C).squeeze(1)
    return x.permute(0, 2, 1)


def dit_forward(sd: SD, inp: torch.Tensor, t: torch.Tensor, text: Optional[torch.Tensor],
                return_hidden: bool = False):
    """Transformer.forward, transformer.py:158-193."""
    x = dit_embed(sd, inp)
    c = time_embedding(t)
    if text is not None:
        c = c + text                                           # transformer.py:176-178
    hidden = [x]
    for l in range(N_LAYERS):
        x = dit_layer(sd, l, x, c)
        hidden.append(x)
    out = dit_unembed(sd, x)
    return (out, hidden) if return_hidden else out


# ----------------------------------------------------------------------------------------------
# Rectified flow (model/backbone/rectified_flow.py) and DDPM (model/backbone/DDPM.py)
# ----------------------------------------------------------------------------------------------

def rf_euler(x_t, v, dt):
    """RectifiedFlow.euler, rectified_flow.py:5-7"""
    return x_t + v * dt


def rf_create_flow(x_1, t, x_0):
    """RectifiedFlow.create_flow, rectified_flow.py:8-12, with the noise x_0 supplied."""
    t = t[:, None, None]
    return t * x_1 + (1 - t) * x_0


def mse(a, b):
    """rectified_flow.py:13-16 / DDPM.py:37-38"""
    return F.mse_loss(a, b)


def ddpm_schedule(total_steps: int):
    """DDPM.__init__, DDPM.py:11-18 -> (beta, alpha, alpha_bar)"""
    beta = torch.linspace(0.0001, 0.02, total_steps)
    alpha = 1 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return beta, alpha, alpha_bar


def _gather(consts, t):
    """DDPM.py:7-9"""
    return consts.gather(-1, t).reshape(-1, 1, 1)


def ddpm_q_sample(x0, t, eps, sched):
    """DDPM.q_xt_x0 + q_sample, DDPM.py:19-27"""
    _, _, alpha_bar = sched
    mean = _gather(alpha_bar, t) ** 0.5 * x0
    var = 1 - _gather(alpha_bar, t)
    return mean + (var ** 0.5) * eps


def ddpm_p_sample(xt, eps_pred, t, noise, sched):
    """DDPM.p_sample, DDPM.py:28-36, with the per-step Gaussian ``noise`` supplied
    (the reference draws torch.randn inside, DDPM.py:35; noise is added at t == 0 too)."""
    beta, alpha, alpha_bar = sched
    ab = _gather(alpha_bar, t)
    a = _gather(alpha, t)
    eps_coef = (1 - a) / (1 - ab) ** .5
    mean = 1 / (a ** 0.5) * (xt - eps_coef * eps_pred)
    var = _gather(beta, t)
    return mean + (var ** .5) * noise


# ----------------------------------------------------------------------------------------------
# LA-VAE (model/pretrained/vqvae.py)
# ----------------------------------------------------------------------------------------------

def _residual_stack(sd: SD, pre: str, x: torch.Tensor, n_layers: int = 2) -> torch.Tensor:
    """ResidualStack.forward vqvae.py:30-33 over Residual.forward vqvae.py:21-22.
    nn.ReLU(True) as the first op of _block mutates x in place before the add, so each layer
    returns relu(x) + conv1(relu(conv3(relu(x))))."""
    for i in range(n_layers):
        r = F.relu(x)
        y = F.conv1d(r, sd[f"{pre}_residual_stack._layers.{i}._block.1.weight"], None, padding=1)
        y = F.relu(y)
        y = F.conv1d(y, sd[f"{pre}_residual_stack._layers.{i}._block.3.weight"], None)
        x = r + y
    return F.relu(x)


def vae_encode(sd: SD, x: torch.Tensor, pre: str = "encoder.") -> Tuple[torch.Tensor, torch.Tensor]:
    """Encoder.forward, vqvae.py:57-71.  x (B,L) -> z (B,64,30), before (B,64,L/4)."""
    x = x.view(x.shape[0], 1, x.shape[-1])
    x = F.relu(F.conv1d(x, sd[pre + "_conv_1.weight"], sd[pre + "_conv_1.bias"], stride=2, padding=1))
    x = F.relu(F.conv1d(x, sd[pre + "_conv_2.weight"], sd[pre + "_conv_2.bias"], stride=2, padding=1))
    x = F.conv1d(x, sd[pre + "_conv_3.weight"], sd[pre + "_conv_3.bias"], padding=1)
    x = _residual_stack(sd, pre, x)
    before = F.conv1d(x, sd[pre + "_pre_vq_conv.weight"], sd[pre + "_pre_vq_conv.bias"])
    z = F.interpolate(before, size=30, mode="linear", align_corners=True)
    return z, before


def vae_decode(sd: SD, z: torch.Tensor, length: int, pre: str = "decoder.") -> Tuple[torch.Tensor, torch.Tensor]:
    """Decoder.forward, vqvae.py:97-105.  z (B,64,30) -> series (B,L) [torch.squeeze'd], after (B,64,L/4)."""
    x = F.interpolate(z, size=int(length / 4), mode="linear", align_corners=True)
    after = x
    x = F.conv1d(x, sd[pre + "_conv_1.weight"], sd[pre + "_conv_1.bias"], padding=1)
    x = _residual_stack(sd, pre, x)
    x = F.relu(F.conv_transpose1d(x, sd[pre + "_conv_trans_1.weight"], sd[pre + "_conv_trans_1.bias"], stride=2, padding=1))
    x = F.conv_transpose1d(x, sd[pre + "_conv_trans_2.weight"], sd[pre + "_conv_trans_2.bias"], stride=2, padding=1)
    return torch.squeeze(x), after


def lavae_encode(sd: SD, x: torch.Tensor, flow_dim: int = 30, pre: str = "encoder.") -> Tuple[torch.Tensor, torch.Tensor]:
    """The fork's multivariate Encoder.forward, myvqvae.py:49-61: x (B,C,L) -> z (B,E,flow_dim), before (B,E,n).
    With C = 1 and flow_dim = 30 this is vqvae.py:57-71."""
    n_layers = sum(1 for k in sd if k.startswith(pre + "_residual_stack._layers.") and k.endswith("_block.1.weight"))
    x = F.relu(F.conv1d(x, sd[pre + "_conv_1.weight"], sd[pre + "_conv_1.bias"], stride=2, padding=1))
    x = F.relu(F.conv1d(x, sd[pre + "_conv_2.weight"], sd[pre + "_conv_2.bias"], stride=2, padding=1))
    x = F.conv1d(x, sd[pre + "_conv_3.weight"], sd[pre + "_conv_3.bias"], padding=1)
    x = _residual_stack(sd, pre, x, n_layers)
    before = F.conv1d(x, sd[pre + "_pre_vq_conv.weight"], sd[pre + "_pre_vq_conv.bias"])
    return F.interpolate(before, size=flow_dim, mode="linear", align_corners=True), before


def lavae_decode(sd: SD, z: torch.Tensor, length: int, pre: str = "decoder.") -> Tuple[torch.Tensor, torch.Tensor]:
    """The fork's multivariate Decoder.forward, myvqvae.py:76-86: z (B,E,flow_dim) -> recon (B,C,length), after
    (B,E,length//4); the last interpolation is the identity when 4 (length // 4) == length."""
    n_layers = sum(1 for k in sd if k.startswith(pre + "_residual_stack._layers.") and k.endswith("_block.1.weight"))
    x = F.interpolate(z, size=int(length / 4), mode="linear", align_corners=True)
    after = x
    x = F.conv1d(x, sd[pre + "_conv_1.weight"], sd[pre + "_conv_1.bias"], padding=1)
    x = _residual_stack(sd, pre, x, n_layers)
    x = F.relu(F.conv_transpose1d(x, sd[pre + "_conv_trans_1.weight"], sd[pre + "_conv_trans_1.bias"], stride=2, padding=1))
    x = F.conv_transpose1d(x, sd[pre + "_conv_trans_2.weight"], sd[pre + "_conv_trans_2.bias"], stride=2, padding=1)
    return F.interpolate(x, size=length, mode="linear", align_corners=True), after


def lavae_train_grads(sd: SD, batch: torch.Tensor, flow_dim: int = 30):
    """vqvae.shared_eval 'train' up to loss.backward() (vqvae.py:121-127, myvqvae.py:121-127):
    batch (B,L) [univariate vqvae.py] or (B,C,L) [myvqvae.py] -> (loss, recon_error, recon, z, {name: grad})."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    x = batch.view(batch.shape[0], 1, batch.shape[-1]) if batch.dim() == 2 else batch
    z, before = lavae_encode(p, x, flow_dim)
    recon, after = lavae_decode(p, z, x.shape[-1])
    data_recon = torch.squeeze(recon) if batch.dim() == 2 else recon          # vqvae.py:105
    recon_error = F.mse_loss(data_recon, batch)
    loss = recon_error + F.mse_loss(before, after)
    loss.backward()
    return loss.detach(), recon_error.detach(), data_recon.detach(), z.detach(), {k: v.grad for k, v in p.items()}


# ----------------------------------------------------------------------------------------------
# Sampling loops (infer.py:75-95) and the training step (train.py:66-87)
# ----------------------------------------------------------------------------------------------

def rf_timesteps(steps: int, batch: int) -> torch.Tensor:
    """infer.py:78 — t_j = round(full(j/steps) * steps) / steps, float32, (steps, batch)."""
    return torch.stack([torch.round(torch.full((batch,), j * 1.0 / steps) * steps) / steps for j in range(steps)])


@torch.no_grad()
def rf_sample(dit: SD, vae: Optional[SD], noise: torch.Tensor, emb: torch.Tensor, steps: int,
              cfg_scale: float, length: Optional[int] = None, return_velocities: bool = False):
    """infer.py:75-82 (+ :95 decode when ``vae`` and ``length`` are given); the initial latent
    ``noise`` replaces randn_like (infer.py:75).  Returns (latent, series|None[, velocities])."""
    x_t = noise.clone()
    vel = []
    for j in range(steps):
        t = torch.round(torch.full((x_t.shape[0],), j * 1.0 / steps) * steps) / steps
        u = dit_forward(dit, x_t, t, None)
        c = dit_forward(dit, x_t, t, emb)
        pred = u + cfg_scale * (c - u)
        if return_velocities:
            vel.append(pred)
        x_t = rf_euler(x_t, pred, 1.0 / steps)
    series = vae_decode(vae, x_t, length)[0] if (vae is not None and length) else None
    return (x_t, series, vel) if return_velocities else (x_t, series)


class _Indexable:
    def __init__(self, fn):
        self.fn = fn

    def __getitem__(self, j):
        return self.fn(j)


@torch.no_grad()
def ddpm_sample(dit: SD, vae: Optional[SD], noise: torch.Tensor, emb: torch.Tensor, steps: int,
                cfg_scale: float, step_noise: torch.Tensor, length: Optional[int] = None,
                return_eps: bool = False, keep_states: Optional[dict] = None):
    """infer.py:83-88 (+ :95).  ``step_noise`` (steps,B,64,30) — a tensor, or a callable j -> (B,64,30) for long
    schedules — replaces the torch.randn drawn inside DDPM.p_sample (DDPM.py:35).  ``keep_states``: a dict whose keys are
    loop indices j; it is filled with (x_t entering step j, guided epsilon of step j) for teacher-forced comparisons."""
    sched = ddpm_schedule(steps)
    x_t = noise.clone()
    eps_l = []
    if not callable(step_noise):
        _sn = step_noise
        step_noise = lambda j: _sn[j]
    step_noise = _Indexable(step_noise)
    for j in range(steps):
        t = torch.full((x_t.size(0),), math.floor(steps - 1 - j), dtype=torch.long)
        u = dit_forward(dit, x_t, t, None)
        c = dit_forward(dit, x_t, t, emb)
        pred = u + cfg_scale * (c - u)
        if return_eps:
            eps_l.append(pred)
        if keep_states is not None and j in keep_states:
            keep_states[j] = (x_t.clone(), pred.clone())
        x_t = ddpm_p_sample(x_t, pred, t, step_noise[j], sched)
    series = vae_decode(vae, x_t, length)[0] if (vae is not None and length) else None
    return (x_t, series, eps_l) if return_eps else (x_t, series)


TRAINABLE_EXCLUDE = ("pos_embed", "unpatch.", "encoder.")


def train_step_grads(dit: SD, x_t: torch.Tensor, t: torch.Tensor, emb: Optional[torch.Tensor],
                     target: torch.Tensor):
    """One forward/backward of train.py:83-85: pred = model(x_t,t,emb|None); loss = mse(pred,target).
    Returns (loss, {name: grad}) for every parameter that receives a gradient."""
    names = [k for k in dit if not k.startswith(TRAINABLE_EXCLUDE)]
    leaf = {k: dit[k].detach().clone().requires_grad_(True) for k in names}
    sd = dict(dit)
    sd.update(leaf)
    loss = mse(dit_forward(sd, x_t, t, emb), target)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), dict(zip(names, grads))


def adamw_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.AdamW single-tensor update as configured at train.py:37 (lr 1e-4, wd 0)."""
    p = p * (1 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


# ---------------------------------------------------------------------------------- evaluation of generated series
def calculate_mse(ori_data, gen_data) -> float:
    """evaluation.py:166-181: arrays (n_samples, a, b); per sample the mean over axis 1 of the squared error, averaged
    over the b slices and then over the samples (= the plain mean for equal-sized slices, kept in the reference's order
    of means)."""
    import numpy as np
    ori, gen = np.asarray(ori_data, dtype=np.float64), np.asarray(gen_data, dtype=np.float64)
    per_slice = ((ori - gen) ** 2).mean(axis=1)          # (n, b): np.mean over [i, :, j]
    return float(per_slice.mean(axis=1).mean())


def calculate_wape(ori_data, gen_data) -> float:
    """evaluation.py:184-206: per sample sum |ori - gen| / sum |ori| over all its elements (NaN when the denominator
    is 0), then np.nanmean over the samples."""
    import numpy as np
    ori, gen = np.asarray(ori_data, dtype=np.float64), np.asarray(gen_data, dtype=np.float64)
    num = np.abs(ori - gen).reshape(ori.shape[0], -1).sum(axis=1)
    den = np.abs(ori).reshape(ori.shape[0], -1).sum(axis=1)
    ratio = np.where(den != 0, num / np.where(den != 0, den, 1.0), np.nan)
    return float(np.nanmean(ratio)) if np.any(den != 0) else float("nan")


# ---------------------------------------------------------------------------------- counter-based DDPM noise
def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123):
    counter words c0..c3 (uint64 arrays holding 32-bit values), key (k0, k1) -> four 32-bit output words.  Pinned by the
    Random123 known-answer vectors in tests/test_oracle_golden.py."""
    import numpy as np
    m32 = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & m32, p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    return c0, c1, c2, c3


def philox_normal(seed: int, step: int, n_elements: int):
    """numpy restatement of the in-kernel noise of t2s_sample_ddpm_seeded (csrc/common.cuh: philox_normal): Philox4x32-10
    keyed by `seed` (low word, high word) on the counter (element, element >> 32, step, 0); u1 = ((w0 >> 8) + 1) 2^-24,
    u2 = (w1 >> 8) 2^-24; z = sqrt(-2 ln u1) cos(2 pi u2).  Returns float32 [n_elements].  It stands in for the torch.randn
    drawn inside DDPM.p_sample (DDPM.py:35), which no implementation can reproduce bit for bit on another generator."""
    import numpy as np
    el = np.arange(n_elements, dtype=np.uint64)
    w0, w1, _, _ = philox4x32_10(el & np.uint64(0xFFFFFFFF), el >> np.uint64(32), np.full(n_elements, step, dtype=np.uint64),
                                 np.zeros(n_elements, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = ((w0 >> np.uint64(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (w1 >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    return (r * np.cos(2.0 * np.pi * u2.astype(np.float64)).astype(np.float32)).astype(np.float32)
