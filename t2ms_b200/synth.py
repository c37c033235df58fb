"""Deterministic synthetic weights and inputs for parity tests and the benchmark.

There is no network for checkpoints, so every measurement uses random-init weights of the
reference architecture.  The distributions follow the reference initialisers:

* DiT (model/denoiser/transformer.py:194-204): xavier-uniform Linear weights, zero Linear biases,
  PyTorch-default Conv2d / ConvTranspose2d / LayerNorm init.  The reference zero-initialises
  ``adaLN_modulation[-1]`` which turns every block into the identity, so here it is
  drawn from N(0, 0.02^2) to exercise attention, the MLP, the text conditioning and CFG.
* LA-VAE (model/pretrained/vqvae.py:36-105): PyTorch-default Conv1d / ConvTranspose1d init,
  hyper-parameters 128/2/256/64 (pretrained_lavae_unified.py:119-122).

Everything is generated on the CPU from a seeded ``torch.Generator`` so that the same state dict
is reproduced bit-for-bit wherever this image runs; ``state_checksum`` guards the golden fixtures
against RNG drift.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict

import torch

D = 128
N_TOK = 480


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _xavier(g, out_f, in_f):
    return _uniform(g, (out_f, in_f), math.sqrt(6.0 / (in_f + out_f)))


def _conv_default(g, shape, fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return _uniform(g, shape, b)


def pos_embed(num_positions: int = N_TOK, d_model: int = D) -> torch.Tensor:
    """Sinusoidal table of model/denoiser/transformer.py:14-23, (1, N, d)."""
    position = torch.arange(num_positions).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model)).unsqueeze(0)
    pe = torch.zeros(num_positions, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def make_dit_state(seed: int = 0, adaln_std: float = 0.02, bias_std: float = 0.0, dim: int = 30) -> Dict[str, torch.Tensor]:
    """State dict with the 55 key names / shapes of the reference ``Transformer``; ``dim`` = latent width
    H of the fork's ``Transformer(dim)`` (only ``pos_embed`` (1, 16 dim, 128) depends on it; the random draws do not).

    ``bias_std`` > 0 additionally randomises the Linear biases and LayerNorm affine (the reference
    init leaves them 0 / 1) so that parity tests exercise every bias path of the kernels.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def bias(n):
        if bias_std > 0:
            return torch.randn(n, generator=g) * bias_std
        return torch.zeros(n)

    sd["pos_embed"] = pos_embed(16 * dim)
    sd["conv.weight"] = _conv_default(g, (4, 1, 2, 2), 4)
    sd["conv.bias"] = _conv_default(g, (4,), 4)
    sd["patch_emb.weight"] = _xavier(g, D, 4)
    sd["patch_emb.bias"] = bias(D)
    sd["ln.weight"] = torch.ones(D) + (torch.randn(D, generator=g) * bias_std if bias_std > 0 else 0)
    sd["ln.bias"] = bias(D)
    sd["linear_emb_to_patch.weight"] = _xavier(g, 4, D)
    sd["linear_emb_to_patch.bias"] = bias(4)
    for l in range(4):
        p = f"layers.{l}."
        sd[p + "attn.qkv.weight"] = _xavier(g, 3 * D, D)
        sd[p + "attn.qkv.bias"] = bias(3 * D)
        sd[p + "attn.proj.weight"] = _xavier(g, D, D)
        sd[p + "attn.proj.bias"] = bias(D)
        sd[p + "mlp.fc1.weight"] = _xavier(g, 2 * D, D)
        sd[p + "mlp.fc1.bias"] = bias(2 * D)
        sd[p + "mlp.fc2.weight"] = _xavier(g, D, 2 * D)
        sd[p + "mlp.fc2.bias"] = bias(D)
        sd[p + "adaLN_modulation.1.weight"] = torch.randn(6 * D, D, generator=g) * adaln_std
        sd[p + "adaLN_modulation.1.bias"] = torch.randn(6 * D, generator=g) * adaln_std
    # unused InverseLatentEmbedding (transformer.py:65-87,150): carried through save/load only
    sd["unpatch.inv_embedding2d.weight"] = _conv_default(g, (D, 1, 6, 6), 36)
    sd["unpatch.inv_embedding2d.bias"] = _conv_default(g, (1,), 36)
    sd["unpatch.fc1.weight"] = _xavier(g, 128, 60)
    sd["unpatch.fc1.bias"] = torch.zeros(128)
    sd["unpatch.fc2.weight"] = _xavier(g, 64, 128)
    sd["unpatch.fc2.bias"] = torch.zeros(64)
    return sd


def make_vae_state(seed: int = 1, hidden: int = 128, res_hidden: int = 256, emb: int = 64,
                   n_res: int = 2, in_channels: int = 1) -> Dict[str, torch.Tensor]:
    """State dict with the key names / shapes of the reference ``vqvae`` (encoder.* / decoder.*); ``in_channels`` > 1
    gives the fork's multivariate ``myvqvae`` (input_dim series channels)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, out_c, in_c, k, bias=True):
        fan_in = in_c * k
        sd[name + ".weight"] = _conv_default(g, (out_c, in_c, k), fan_in)
        if bias:
            sd[name + ".bias"] = _conv_default(g, (out_c,), fan_in)

    def convT(name, in_c, out_c, k):
        fan_in = out_c * k  # torch computes fan_in from weight.size(1) * k for ConvTranspose
        sd[name + ".weight"] = _conv_default(g, (in_c, out_c, k), fan_in)
        sd[name + ".bias"] = _conv_default(g, (out_c,), fan_in)

    conv("encoder._conv_1", hidden // 2, in_channels, 4)
    conv("encoder._conv_2", hidden, hidden // 2, 4)
    conv("encoder._conv_3", hidden, hidden, 3)
    for i in range(n_res):
        conv(f"encoder._residual_stack._layers.{i}._block.1", res_hidden, hidden, 3, bias=False)
        conv(f"encoder._residual_stack._layers.{i}._block.3", hidden, res_hidden, 1, bias=False)
    conv("encoder._pre_vq_conv", emb, hidden, 1)
    conv("decoder._conv_1", hidden, emb, 3)
    for i in range(n_res):
        conv(f"decoder._residual_stack._layers.{i}._block.1", res_hidden, hidden, 3, bias=False)
        conv(f"decoder._residual_stack._layers.{i}._block.3", hidden, res_hidden, 1, bias=False)
    convT("decoder._conv_trans_1", hidden, hidden // 2, 4)
    convT("decoder._conv_trans_2", hidden // 2, in_channels, 4)
    return sd


def make_text_embeddings(batch: int, seed: int = 2) -> torch.Tensor:
    """Unit-norm 128-d vectors (OpenAI text-embedding-3-large @ dimensions=128,
    Dataset_Construction_Pipeline/Get_Embedding_and_Convert_JSON_to_CSV.py:15-17)."""
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(batch, D, generator=g), dim=-1)


def make_noise(batch: int, seed: int = 3, dim: int = 30) -> torch.Tensor:
    """Initial latent replacing randn_like at infer.py:75, (B,64,dim)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 64, dim, generator=g)


def make_step_noise(steps: int, batch: int, seed: int = 4, dim: int = 30) -> torch.Tensor:
    """Pre-drawn Gaussian noise replacing torch.randn inside DDPM.p_sample (DDPM.py:35)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(steps, batch, 64, dim, generator=g)


def make_series(batch: int, length: int, seed: int = 5) -> torch.Tensor:
    """MinMax-scaled series in [0,1] (datafactory/dataset.py:81-82), (B,L)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, length, generator=g)


def state_checksum(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().float().numpy().tobytes())
    return h.hexdigest()
