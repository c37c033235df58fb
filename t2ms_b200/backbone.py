"""Diffusion process maths with the reference interfaces (model/backbone/rectified_flow.py,
model/backbone/DDPM.py).

Inside the fused sampler (t2ms_b200.sampler) the Euler / ancestral updates and the guidance mix run
in the epilogue of the last DiT kernel; these classes keep the reference's step-wise API for
``infer.py`` / ``train.py`` style loops and provide the per-step coefficient tables the fused
kernels consume (computed with the same fp32 torch expressions as the reference).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


class RectifiedFlow:
    """model/backbone/rectified_flow.py:4-16"""

    def euler(self, x_t, v, dt):
        return x_t + v * dt

    def create_flow(self, x_1, t):
        x_0 = torch.randn_like(x_1).to(x_1.device)
        t = t[:, None, None]
        x_t = t * x_1 + (1 - t) * x_0
        return x_t, x_0

    def loss(self, v, noise_gt):
        return F.mse_loss(v, noise_gt)

    # ---- tables for the fused sampler
    @staticmethod
    def timesteps(steps: int) -> torch.Tensor:
        """t_j of infer.py:78: round(full(j/steps) * steps) / steps in fp32, (steps,)."""
        return torch.cat([torch.round(torch.full((1,), j * 1.0 / steps) * steps) / steps for j in range(steps)])

    @staticmethod
    def coefficients(steps: int) -> torch.Tensor:
        """(steps,3) fp32: {dt, 0, 0} with dt = 1.0/steps (infer.py:82)."""
        c = torch.zeros(steps, 3, dtype=torch.float32)
        c[:, 0] = torch.tensor(1.0 / steps, dtype=torch.float32)
        return c


def gather(consts: torch.Tensor, t: torch.Tensor):
    """model/backbone/DDPM.py:7-9"""
    c = consts.gather(-1, t)
    return c.reshape(-1, 1, 1)


class DDPM:
    """model/backbone/DDPM.py:10-38"""

    def __init__(self, total_steps: int, device):
        self.device = device
        self.beta = torch.linspace(0.0001, 0.02, total_steps).to(device)
        self.alpha = 1 - self.beta
        self.alpha_bar = torch.cumprod(self.alpha, dim=0)
        self.total_steps = total_steps
        self.sigma2 = self.beta

    def q_xt_x0(self, x0: torch.Tensor, t: torch.Tensor):
        mean = gather(self.alpha_bar, t) ** 0.5 * x0
        var = 1 - gather(self.alpha_bar, t)
        return mean.to(self.device), var.to(self.device)

    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, eps: Optional[torch.Tensor] = None):
        if eps is None:
            eps = torch.randn_like(x0).to(self.device)
        mean, var = self.q_xt_x0(x0, t)
        return (mean + (var ** 0.5) * eps).to(self.device), eps

    def p_sample(self, xt: torch.Tensor, n_xt: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        alpha_bar = gather(self.alpha_bar, t)
        alpha = gather(self.alpha, t)
        eps_coef = (1 - alpha) / (1 - alpha_bar) ** .5
        mean = 1 / (alpha ** 0.5) * (xt - eps_coef * n_xt)
        var = gather(self.sigma2, t)
        eps = torch.randn(xt.shape, device=xt.device)
        return mean + (var ** .5) * eps

    def loss(self, n_gt: torch.Tensor, n_xt: torch.Tensor):
        return F.mse_loss(n_gt, n_xt)

    # ---- tables for the fused sampler
    @staticmethod
    def timesteps(steps: int) -> torch.Tensor:
        """t_j = floor(steps-1-j) of infer.py:84, as the float the time embedding sees, (steps,)."""
        return torch.tensor([math.floor(steps - 1 - j) for j in range(steps)], dtype=torch.long)

    @staticmethod
    def coefficients(steps: int) -> torch.Tensor:
        """(steps,3) fp32 rows {1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-alpha_bar_t), sqrt(beta_t)} for
        t = steps-1-j, evaluated with the expressions of DDPM.py:14-18,30-34 on the CPU."""
        beta = torch.linspace(0.0001, 0.02, steps)
        alpha = 1 - beta
        alpha_bar = torch.cumprod(alpha, dim=0)
        t = DDPM.timesteps(steps)
        a, ab, var = alpha[t], alpha_bar[t], beta[t]
        eps_coef = (1 - a) / (1 - ab) ** .5
        return torch.stack([1 / (a ** 0.5), eps_coef, var ** .5], dim=1).to(torch.float32).contiguous()
