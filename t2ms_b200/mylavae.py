"""The fork's multivariate LA-VAE with the reference module interface (model/pretrained/myvqvae.py): series
(B, input_dim, L) of any length, latent (B, embedding_dim, flow_dim).  Same parameter names / shapes as the reference
(``encoder._conv_1.weight`` ... ``decoder._conv_trans_2.bias``), so its state dicts load with ``strict=True``.
Forward, loss and backward run layer-wise through the C ABI (t2s_lavae_encode / _decode / _train_step); CUDA only.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .lavae import BaseModel, ResidualStack
from .lavae_train import LavaeEngine


class Encoder(nn.Module):
    """model/pretrained/myvqvae.py:32-61 (parameters; forward = t2s_lavae_encode)."""

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens, embedding_dim, flow_dim):
        super().__init__()
        self.flow_dim = flow_dim
        self._conv_1 = nn.Conv1d(in_channels, num_hiddens // 2, kernel_size=4, stride=2, padding=1)
        self._conv_2 = nn.Conv1d(num_hiddens // 2, num_hiddens, kernel_size=4, stride=2, padding=1)
        self._conv_3 = nn.Conv1d(num_hiddens, num_hiddens, kernel_size=3, stride=1, padding=1)
        self._residual_stack = ResidualStack(num_hiddens, num_hiddens, num_residual_layers, num_residual_hiddens)
        self._pre_vq_conv = nn.Conv1d(num_hiddens, embedding_dim, kernel_size=1, stride=1)

    def forward(self, inputs):
        return self._owner_engine().encode(inputs)


class Decoder(nn.Module):
    """model/pretrained/myvqvae.py:63-86 (parameters; forward = t2s_lavae_decode, incl. the final interpolation :85)."""

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens, out_channels=52):
        super().__init__()
        self._conv_1 = nn.Conv1d(in_channels, num_hiddens, kernel_size=3, stride=1, padding=1)
        self._residual_stack = ResidualStack(num_hiddens, num_hiddens, num_residual_layers, num_residual_hiddens)
        self._conv_trans_1 = nn.ConvTranspose1d(num_hiddens, num_hiddens // 2, kernel_size=4, stride=2, padding=1)
        self._conv_trans_2 = nn.ConvTranspose1d(num_hiddens // 2, out_channels, kernel_size=4, stride=2, padding=1)

    def forward(self, inputs, length):
        return self._owner_engine().decode(inputs, length)


class vqvae(BaseModel):
    """model/pretrained/myvqvae.py:88-156"""

    def __init__(self, args):
        super().__init__()
        self.encoder = Encoder(args.input_dim, args.block_hidden_size, args.num_residual_layers, args.res_hidden_size,
                               args.embedding_dim, args.flow_dim)
        self.decoder = Decoder(args.embedding_dim, args.block_hidden_size, args.num_residual_layers, args.res_hidden_size,
                               out_channels=args.input_dim)
        self._engine = LavaeEngine(self, flow_dim=args.flow_dim)
        # the sub-modules reach the engine (which needs both halves' parameters for the struct) through a closure,
        # not a registered attribute, so no module cycle is created
        eng = self._engine
        object.__setattr__(self.encoder, "_owner_engine", lambda: eng)
        object.__setattr__(self.decoder, "_owner_engine", lambda: eng)

    def shared_eval(self, batch, optimizer, mode):
        """myvqvae.py:116-136"""
        if mode == "train":
            optimizer.zero_grad()
            loss, recon_error, recon, z = self._engine.step(batch, backward=True)
            optimizer.step()
        else:
            loss, recon_error, recon, z = self._engine.step(batch, backward=False)
        return loss, recon_error, recon, z

    def forward(self, x):
        """myvqvae.py:138-142"""
        z, _ = self.encoder(x)
        out, _ = self.decoder(z, length=x.shape[-1])
        return out

    def custom_loss(self, x, y, lambda_smooth=0.1):
        """myvqvae.py:144-156 (plain torch: not on any measured path)."""
        smooth_l1_loss = F.smooth_l1_loss(x, y)
        x_diff, y_diff = x[..., 1:] - x[..., :-1], y[..., 1:] - y[..., :-1]
        return smooth_l1_loss + lambda_smooth * F.smooth_l1_loss(x_diff, y_diff)
