"""LA-VAE with the reference module interface (model/pretrained/vqvae.py, core.py).

``vqvae(args).encoder(x) -> (z, before)`` and ``.decoder(z, length) -> (squeeze(series), after)`` keep
the reference's names, shapes and state-dict keys; both forwards are single fused sm_100a kernels
(t2s_vae_encode / t2s_vae_decode in include/t2s_b200.h).  CUDA only.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
import torch.nn as nn
from torch.optim.lr_scheduler import CosineAnnealingLR, LinearLR, SequentialLR

from . import _lib
from .packing import PackedVaeDecoder, PackedVaeEncoder

LENGTHS = (24, 48, 96)


class BaseModel(nn.Module, ABC):
    """model/pretrained/core.py:8-20"""

    def __init__(self):
        super().__init__()

    @abstractmethod
    def shared_eval(self, batch, optimizer, scheduler, mode):
        pass

    def configure_optimizers(self, lr=1e-3):
        optimizer = torch.optim.AdamW(self.parameters(), lr=lr, weight_decay=1e-2)
        scheduler1 = LinearLR(optimizer, start_factor=0.1, total_iters=1000)
        scheduler2 = CosineAnnealingLR(optimizer, T_max=400 - 1000, eta_min=1e-6)
        scheduler = SequentialLR(optimizer, schedulers=[scheduler1, scheduler2], milestones=[1000])
        return optimizer, scheduler


class Residual(nn.Module):
    """model/pretrained/vqvae.py:7-22 (parameters only; the in-place-ReLU semantics live in the kernel)."""

    def __init__(self, in_channels, num_hiddens, num_residual_hiddens):
        super().__init__()
        self._block = nn.Sequential(
            nn.ReLU(True),
            nn.Conv1d(in_channels, num_residual_hiddens, kernel_size=3, stride=1, padding=1, bias=False),
            nn.ReLU(True),
            nn.Conv1d(num_residual_hiddens, num_hiddens, kernel_size=1, stride=1, bias=False))


class ResidualStack(nn.Module):
    """model/pretrained/vqvae.py:24-33"""

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens):
        super().__init__()
        self._num_residual_layers = num_residual_layers
        self._layers = nn.ModuleList([Residual(in_channels, num_hiddens, num_residual_hiddens)
                                      for _ in range(num_residual_layers)])


class _PackedMixin:
    _packer = None

    def _packed_weights(self):
        params = list(self.named_parameters())
        dev = params[0][1].device
        if dev.type != "cuda":
            raise RuntimeError(f"t2ms_b200 {type(self).__name__} runs on CUDA (sm_100a) only (no CPU fallback)")
        key = (str(dev), tuple(p._version for _, p in params), tuple(p.data_ptr() for _, p in params))
        if getattr(self, "_pk", None) is None or self._pk_key != key:
            with torch.no_grad():
                self._pk = type(self)._packer({n: p for n, p in params}, dev)
            self._pk_key = key
            self._pk_generation = getattr(self, "_pk_generation", 0) + 1
            self._pk.generation = self._pk_generation
        return self._pk

    def _warn_if_grad_expected(self, inputs):
        """The fused single-kernel encoder / decoder forwards carry no autograd graph (they serve inference and the frozen
        encoder of train.py:30-33).  Say so once instead of silently returning detached tensors when the caller seems to
        expect gradients; LA-VAE training goes through ``vqvae.shared_eval`` (t2s_lavae_train_step)."""
        if torch.is_grad_enabled() and not getattr(self, "_warned_no_grad", False) and \
                (inputs.requires_grad or any(p.requires_grad for p in self.parameters())):
            import warnings
            self._warned_no_grad = True
            warnings.warn(f"t2ms_b200 {type(self).__name__}.forward returns tensors WITHOUT a grad_fn (fused inference kernel): freeze the "
                          "module / use torch.no_grad() as train.py:31-33 and infer.py:65 do, or train the LA-VAE with vqvae.shared_eval.",
                          stacklevel=3)

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_pk", None)
        st.pop("_pk_key", None)
        return st


def _check_arch(num_hiddens, num_residual_layers, num_residual_hiddens, embedding_dim):
    if (num_hiddens, num_residual_layers, num_residual_hiddens, embedding_dim) != (128, 2, 256, 64):
        raise ValueError("t2ms_b200 LA-VAE kernels are specialised for block_hidden_size=128, num_residual_layers=2, "
                         "res_hidden_size=256, embedding_dim=64 (pretrained_lavae_unified.py:119-122)")


class Encoder(_PackedMixin, nn.Module):
    """model/pretrained/vqvae.py:36-71"""
    _packer = PackedVaeEncoder

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens, embedding_dim):
        super().__init__()
        if in_channels != 1:
            raise ValueError("univariate series only (in_channels=1)")
        _check_arch(num_hiddens, num_residual_layers, num_residual_hiddens, embedding_dim)
        self._conv_1 = nn.Conv1d(in_channels, num_hiddens // 2, kernel_size=4, stride=2, padding=1)
        self._conv_2 = nn.Conv1d(num_hiddens // 2, num_hiddens, kernel_size=4, stride=2, padding=1)
        self._conv_3 = nn.Conv1d(num_hiddens, num_hiddens, kernel_size=3, stride=1, padding=1)
        self._residual_stack = ResidualStack(num_hiddens, num_hiddens, num_residual_layers, num_residual_hiddens)
        self._pre_vq_conv = nn.Conv1d(num_hiddens, embedding_dim, kernel_size=1, stride=1)

    def forward(self, inputs):
        if not inputs.is_cuda:
            raise RuntimeError("t2ms_b200 Encoder.forward needs CUDA tensors (no CPU fallback)")
        self._warn_if_grad_expected(inputs)
        B, L = inputs.shape[0], inputs.shape[-1]
        if L not in LENGTHS:
            raise ValueError(f"series length must be one of {LENGTHS}, got {L}")
        from . import ops
        x = inputs.detach().reshape(B, L).to(torch.float32).contiguous()
        return ops.vae_encode(x, self._packed_weights().handle)          # custom op t2s_b200::vae_encode -> t2s_vae_encode


class Decoder(_PackedMixin, nn.Module):
    """model/pretrained/vqvae.py:74-105"""
    _packer = PackedVaeDecoder

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens):
        super().__init__()
        _check_arch(num_hiddens, num_residual_layers, num_residual_hiddens, in_channels)
        self._conv_1 = nn.Conv1d(in_channels, num_hiddens, kernel_size=3, stride=1, padding=1)
        self._residual_stack = ResidualStack(num_hiddens, num_hiddens, num_residual_layers, num_residual_hiddens)
        self._conv_trans_1 = nn.ConvTranspose1d(num_hiddens, num_hiddens // 2, kernel_size=4, stride=2, padding=1)
        self._conv_trans_2 = nn.ConvTranspose1d(num_hiddens // 2, 1, kernel_size=4, stride=2, padding=1)

    def decode_into(self, z: torch.Tensor, length: int, series: torch.Tensor, after=None):
        """Decode into a caller-owned (B, length) buffer (custom op t2s_b200::vae_decode_into)."""
        from . import ops
        assert after is None and series.shape[1] == int(length)
        ops.vae_decode_into(z, series, self._packed_weights().handle)

    def forward(self, inputs, length):
        if not inputs.is_cuda:
            raise RuntimeError("t2ms_b200 Decoder.forward needs CUDA tensors (no CPU fallback)")
        self._warn_if_grad_expected(inputs)
        length = int(length)
        if length not in LENGTHS:
            raise ValueError(f"series length must be one of {LENGTHS}, got {length}")
        z = inputs.detach().to(torch.float32).contiguous()
        assert z.shape[1:] == (64, 30), f"latent must be (B,64,30), got {tuple(z.shape)}"
        from . import ops
        series, after = ops.vae_decode(z, length, self._packed_weights().handle)      # custom op -> t2s_vae_decode
        return torch.squeeze(series.unsqueeze(1)), after          # vqvae.py:105 squeezes (B,1,L): drops the batch dim too at B == 1


class vqvae(BaseModel):
    """model/pretrained/vqvae.py:108-142"""

    def __init__(self, args):
        super().__init__()
        self.encoder = Encoder(1, args.block_hidden_size, args.num_residual_layers, args.res_hidden_size, args.embedding_dim)
        self.decoder = Decoder(args.embedding_dim, args.block_hidden_size, args.num_residual_layers, args.res_hidden_size)

    def shared_eval(self, batch, optimizer, mode):
        """vqvae.py:118-135: encoder, decoder, recon_error = mse(recon, batch), cross_loss = mse(before, after),
        loss = their sum; in 'train' mode ``optimizer.zero_grad()``, the backward (gradients written to ``param.grad``
        by t2s_lavae_train_step) and ``optimizer.step()`` with the caller's optimizer.  Returns
        (loss, recon_error, data_recon, z) like the reference (data_recon is torch.squeeze'd, vqvae.py:105)."""
        from .lavae_train import LavaeEngine
        if getattr(self, "_engine", None) is None:
            self._engine = LavaeEngine(self, flow_dim=30)
        if mode == "train":
            optimizer.zero_grad()
            loss, recon_error, recon, z = self._engine.step(batch, backward=True)
            optimizer.step()
        elif mode in ("val", "test"):
            loss, recon_error, recon, z = self._engine.step(batch, backward=False)
        else:
            raise ValueError(f"unknown mode {mode!r}")
        return loss, recon_error, torch.squeeze(recon), z

    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_engine", None)
        return st

    def forward(self, x):
        raise NotImplementedError("vqvae.forward is broken in the reference (vqvae.py:137-142 passes a tuple "
                                  "to the decoder) and unused; call .encoder / .decoder")
