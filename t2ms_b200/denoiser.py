"""T2S-DiT denoiser with the reference module interface (model/denoiser/transformer.py).

``Transformer`` keeps the reference's constructor, ``forward(input, t, text_input)`` signature,
attribute names and state-dict keys (55 keys) so that ``infer.py`` / ``train.py`` can use
it unchanged; the forward itself is the hand-written sm_100a kernel chain reached through the C ABI
(include/t2s_b200.h: t2s_dit_forward).  There is no PyTorch / CPU fallback: a CPU tensor raises.

The submodules below are parameter containers only (their own ``forward`` is never called).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .packing import PackedDit
from .synth import pos_embed as _pos_embed

D_MODEL = 128


def modulate(x, shift, scale):
    """model/denoiser/transformer.py:7-8 (kept for API parity; fused into the kernels)."""
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def get_sinusoidal_positional_embeddings(num_positions, d_model):
    """model/denoiser/transformer.py:14-23"""
    return _pos_embed(num_positions, d_model)


class TimeEmbedding(nn.Module):
    """model/denoiser/transformer.py:25-40.  Parameter-free; evaluated inside cond_kernel."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        assert dim % 2 == 0, "Dimension must be even"


class _Attention(nn.Module):
    """Parameter container with timm==1.0.11 Attention's names (qkv, proj)."""

    def __init__(self, dim, num_heads=4, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    """Parameter container with timm==1.0.11 Mlp's names (fc1, fc2)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class InverseLatentEmbedding(nn.Module):
    """model/denoiser/transformer.py:65-87 — unused by forward, carried for checkpoint compatibility."""

    def __init__(self, embed_dim: int = 64):
        super().__init__()
        self.dim = embed_dim
        self.inv_embedding2d = nn.ConvTranspose2d(embed_dim, 1, kernel_size=(6, 6), stride=(6, 6))
        self.fc1 = nn.Linear(60, 128)
        self.fc2 = nn.Linear(128, 64)


class Transformerlayer(nn.Module):
    """model/denoiser/transformer.py:94-109 (parameters only)."""

    def __init__(self):
        super().__init__()
        d_model = D_MODEL
        self.norm1 = nn.LayerNorm(d_model, elementwise_affine=False, eps=1e-6)
        self.norm2 = nn.LayerNorm(d_model, elementwise_affine=False, eps=1e-6)
        self.attn = _Attention(d_model, num_heads=4, qkv_bias=True)
        self.mlp = _Mlp(d_model, int(d_model * 2.0))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(d_model, 6 * d_model, bias=True))


class Transformer(nn.Module):
    """Drop-in for model/denoiser/transformer.py:127-204 (``Transformer()``: H = 30, 480 tokens) and for the fork's
    model/denoiser/mytransformer.py:127-204 (``Transformer(dim)``: H = dim; the kernels are built for dim 30, 50 and
    64 = config.yaml:46,91 ``flow_dim``, i.e. 480 / 800 / 1024 tokens).  Latents are (B, 64, H)."""

    def __init__(self, dim: int = 30):
        super().__init__()
        if dim not in (30, 50, 64):
            raise ValueError("t2ms_b200.Transformer: latent width (dim) must be 30 (T2S), 50 or 64 (fork configs)")
        self.channel = 1
        self.H = int(dim)
        self.W = 64
        emb_size = D_MODEL
        self.patch_size = 2
        self.patch_count = int((self.H / self.patch_size) * (self.W / self.patch_size))
        self.conv = nn.Conv2d(self.channel, self.channel * self.patch_size ** 2, kernel_size=self.patch_size,
                              padding=0, stride=self.patch_size)
        self.patch_emb = nn.Linear(self.channel * self.patch_size ** 2, emb_size)
        self.pos_embed = nn.Parameter(get_sinusoidal_positional_embeddings(self.patch_count, emb_size), requires_grad=False)
        self.ln = nn.LayerNorm(emb_size)
        self.linear_emb_to_patch = nn.Linear(emb_size, self.channel * self.patch_size ** 2)
        self.time_emb = TimeEmbedding(dim=emb_size)
        self.layers = nn.ModuleList([Transformerlayer() for _ in range(4)])
        self.unpatch = InverseLatentEmbedding(embed_dim=emb_size)
        self.initialize_weights()
        self._packed: Optional[PackedDit] = None
        self._packed_key = None
        self._pack_generation = 0          # bumped by every re-pack: cache keys use it, never id() of a freed object
        self._workspaces = {}
        self._warned_no_grad = False

    def initialize_weights(self):
        """model/denoiser/transformer.py:194-204 (xavier Linear weights, zero biases, zero adaLN)."""
        def _basic_init(module):
            if isinstance(module, nn.Linear):
                torch.nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.constant_(module.bias, 0)
        self.apply(_basic_init)
        for block in self.layers:
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)

    # ------------------------------------------------------------------ packed weights / workspace
    def _own_params(self):
        return [(n, p) for n, p in self.named_parameters() if not n.startswith("encoder.")]

    def packed(self) -> PackedDit:
        params = self._own_params()
        dev = params[0][1].device
        key = (str(dev), tuple(p._version for _, p in params), tuple(p.data_ptr() for _, p in params))
        if self._packed is None or self._packed_key != key:
            if dev.type != "cuda":
                raise RuntimeError("t2ms_b200.Transformer runs on CUDA (sm_100a) only; move it with .to('cuda')")
            with torch.no_grad():
                self._packed = PackedDit({n: p for n, p in params}, dev)
            self._packed_key = key
            self._pack_generation += 1
            self._packed.generation = self._pack_generation
        return self._packed

    def workspace(self, nseq: int, device) -> torch.Tensor:
        key = (str(device), nseq)
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = _lib.load().t2s_dit_workspace_bytes_h(nseq, self.H)
            ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=device)
            # small LRU; an evicted buffer is only freed once nothing else (a cached CUDA graph of T2SSampler) holds it
            while len(self._workspaces) >= 4:
                self._workspaces.pop(next(iter(self._workspaces)))
            self._workspaces[key] = ws
        else:
            self._workspaces[key] = self._workspaces.pop(key)           # most recently used last
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, input: torch.Tensor, t: torch.Tensor, text_input):
        """input (B,64,H), t (B,) float32 or int64, text_input (B,128) or None -> (B,64,H)."""
        if not input.is_cuda:
            raise RuntimeError("t2ms_b200.Transformer.forward needs CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for _, p in self._own_params()) and self.training:
            from .training import dit_forward_autograd
            return dit_forward_autograd(self, input, t, text_input)
        if torch.is_grad_enabled() and not self._warned_no_grad and (input.requires_grad or (text_input is not None and text_input.requires_grad)
                                                                    or any(p.requires_grad for _, p in self._own_params())):
            import warnings
            self._warned_no_grad = True
            warnings.warn("t2ms_b200.Transformer.forward in eval() mode (or with frozen parameters) returns a tensor WITHOUT a grad_fn; "
                          "gradients w.r.t. `input` / `text_input` are never produced.  Call .train() for the training path, or wrap "
                          "inference in torch.no_grad() as infer.py:65 does.", stacklevel=2)
        return dit_forward(self, input, t, text_input)


def _aligned(ws: torch.Tensor) -> int:
    return (ws.data_ptr() + 255) & ~255


def dit_forward(model: Transformer, x: torch.Tensor, t: torch.Tensor, text: Optional[torch.Tensor]) -> torch.Tensor:
    """Transformer.forward through the custom op ``t2s_b200::dit_forward`` (t2ms_b200/ops.py -> t2s_dit_forward)."""
    from . import ops
    B = x.shape[0]
    assert x.shape[1:] == (64, model.H), f"latent must be (B,64,{model.H}), got {tuple(x.shape)}"
    x = x.detach().to(torch.float32).contiguous()
    t100 = (t.detach() * 100.0).to(torch.float32).contiguous()          # transformer.py:31
    assert t100.shape == (B,)
    if text is not None:
        text = text.detach().to(torch.float32).contiguous()
        assert text.shape == (B, D_MODEL)
    pk = model.packed()
    return ops.dit_forward(x, t100, text, model.workspace(B, x.device), pk.handle, model.H)
