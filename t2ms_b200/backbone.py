"""Diffusion-process classes with the reference's interfaces (model/backbone/rectified_flow.py:4-16,
model/backbone/DDPM.py:7-38), computed by the C-ABI kernels.

Inside the fused sampler (t2ms_b200.sampler) the Euler / ancestral update and the guidance mix run in the epilogue of the
last DiT kernel.  These classes serve loops that call the process once per step, like an unmodified ``infer.py`` /
``train.py``: every method takes CUDA tensors and goes through a ``t2s_b200::*`` custom op (t2ms_b200/ops.py) to a device
kernel that rounds each operation like the reference's torch expression.  CPU tensors raise (no fallback); the schedule
tables are small host-side set-up, built with the reference's formulas.  The static ``timesteps`` / ``coefficients``
helpers give the fused kernels their per-step launch parameters.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("t2ms_b200 backbone classes run on CUDA tensors (sm_100a kernels); there is no CPU fallback")


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _per_sample(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """(B,) coefficient tensor on `like`'s device, fp32 contiguous (one value per sample, broadcast over (64, H) by the kernel)."""
    return t.reshape(-1).to(device=like.device, dtype=torch.float32).contiguous()


class RectifiedFlow:
    """model/backbone/rectified_flow.py:4-16"""

    def euler(self, x_t, v, dt):
        """x_t + v * dt (rectified_flow.py:5-7) -> t2s_rf_euler."""
        from . import ops
        _need_cuda(x_t, v)
        return ops.rf_euler(_f32(x_t), _f32(v), float(dt)).view_as(x_t)

    def create_flow(self, x_1, t):
        """x_0 ~ N(0, I); x_t = t x_1 + (1 - t) x_0 (rectified_flow.py:8-12) -> t2s_train_make_inputs; returns (x_t, x_0)."""
        from . import ops
        _need_cuda(x_1)
        x1 = _f32(x_1)
        x_0 = torch.randn_like(x1)
        x_t, _ = ops.make_inputs(0, x1, x_0, _per_sample(t, x1), None, x1.shape[-1])
        return x_t, x_0

    def loss(self, v, noise_gt):
        """MSE (rectified_flow.py:13-16); stays a torch op so that callers can differentiate through it."""
        return F.mse_loss(v, noise_gt)

    # ---- launch parameters of the fused sampler
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
    """model/backbone/DDPM.py:7-9: one schedule constant per sample, shaped (B,1,1)."""
    return consts.gather(-1, t).reshape(-1, 1, 1)


class DDPM:
    """model/backbone/DDPM.py:10-38: linear beta schedule 1e-4 .. 0.02 over total_steps, sigma^2 = beta."""

    def __init__(self, total_steps: int, device):
        self.device, self.total_steps = device, total_steps
        self.beta = torch.linspace(0.0001, 0.02, total_steps).to(device)
        self.alpha = 1 - self.beta
        self.alpha_bar = torch.cumprod(self.alpha, dim=0)
        self.sigma2 = self.beta

    def q_xt_x0(self, x0: torch.Tensor, t: torch.Tensor):
        """Mean and variance of q(x_t | x_0) (DDPM.py:19-22)."""
        _need_cuda(x0)
        ab = gather(self.alpha_bar.to(x0.device), t.to(x0.device))
        return (ab ** 0.5 * x0).to(self.device), (1 - ab).to(self.device)

    def q_sample(self, x0: torch.Tensor, t: torch.Tensor, eps: Optional[torch.Tensor] = None):
        """sqrt(alpha_bar_t) x0 + sqrt(1 - alpha_bar_t) eps (DDPM.py:23-27) -> t2s_train_make_inputs; returns (x_t, eps)."""
        from . import ops
        _need_cuda(x0, eps)
        x = _f32(x0)
        if eps is None:
            eps = torch.randn_like(x)
        ab = self.alpha_bar.to(x.device).gather(-1, t.to(x.device))
        x_t, _ = ops.make_inputs(1, x, _f32(eps), _per_sample(ab ** 0.5, x), _per_sample((1 - ab) ** 0.5, x), x.shape[-1])
        return x_t.to(self.device), eps

    def p_sample(self, xt: torch.Tensor, n_xt: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Ancestral step (DDPM.py:28-36): (x_t - (1-alpha_t)/sqrt(1-alpha_bar_t) eps_hat)/sqrt(alpha_t) + sqrt(beta_t) z with
        z = torch.randn drawn here, at every step including t = 0, like the reference -> t2s_ddpm_p_sample."""
        from . import ops
        _need_cuda(xt, n_xt)
        x, t = _f32(xt), t.to(xt.device)
        a, ab = self.alpha.to(x.device).gather(-1, t), self.alpha_bar.to(x.device).gather(-1, t)
        c1, c2, c3 = 1 / (a ** 0.5), (1 - a) / (1 - ab) ** .5, self.sigma2.to(x.device).gather(-1, t) ** .5
        z = torch.randn(xt.shape, device=xt.device)
        return ops.ddpm_p_sample(x, _f32(n_xt), z, _per_sample(c1, x), _per_sample(c2, x), _per_sample(c3, x)).view_as(xt)

    def loss(self, n_gt: torch.Tensor, n_xt: torch.Tensor):
        """MSE (DDPM.py:37-38)."""
        return F.mse_loss(n_gt, n_xt)

    # ---- launch parameters of the fused sampler
    @staticmethod
    def timesteps(steps: int) -> torch.Tensor:
        """t_j = floor(steps-1-j) of infer.py:84, (steps,) int64."""
        return torch.tensor([math.floor(steps - 1 - j) for j in range(steps)], dtype=torch.long)

    @staticmethod
    def coefficients(steps: int) -> torch.Tensor:
        """(steps,3) fp32 rows {1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-alpha_bar_t), sqrt(beta_t)} for t = steps-1-j, from the
        schedule of DDPM.py:14-18 evaluated on the CPU."""
        sched = DDPM(steps, "cpu")
        t = DDPM.timesteps(steps)
        a, ab, var = sched.alpha[t], sched.alpha_bar[t], sched.sigma2[t]
        return torch.stack([1 / (a ** 0.5), (1 - a) / (1 - ab) ** .5, var ** .5], dim=1).to(torch.float32).contiguous()
