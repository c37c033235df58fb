"""PyTorch custom-op layer over the C ABI (include/t2s_b200.h).

``BASELINE.json:north_star`` asks for "a thin C-ABI PyTorch custom-op layer": every entry point the module classes use is
registered here with ``torch.library.custom_op`` (namespace ``t2s_b200``) together with a fake (meta) implementation, so the
calls are visible to the dispatcher, traceable by ``torch.compile`` / ``torch.export`` with fake tensors, and safe under
CUDA-graph capture (each real implementation only enqueues kernels on the current stream through ctypes).  The product path
is CUDA only: a CPU tensor raises, there is no eager / PyTorch fallback.

Packed weights are not tensors (they are C structs of device pointers owned by Python objects), so ops take an integer
``handle`` obtained from ``register_handle``; the registry holds weak references.
"""
from __future__ import annotations

import ctypes as C
import itertools
import weakref
from typing import Optional, Tuple

import torch

from . import _lib

_HANDLES: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_NEXT = itertools.count(1)


def register_handle(obj) -> int:
    """Weakly register a packed-weights object (anything with a ``.ref`` ctypes reference); returns its op handle."""
    h = next(_NEXT)
    _HANDLES[h] = obj
    return h


def _obj(handle: int):
    try:
        return _HANDLES[handle]
    except KeyError:
        raise RuntimeError(f"t2s_b200: stale weight handle {handle} (the packed weights were released)") from None


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("t2s_b200 ops need CUDA tensors (sm_100a); there is no CPU / PyTorch fallback")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _aligned(ws: torch.Tensor) -> int:
    return (ws.data_ptr() + 255) & ~255


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


# ---------------------------------------------------------------------------------------------- denoiser forward
@torch.library.custom_op("t2s_b200::dit_forward", mutates_args=("workspace",), device_types="cuda")
def dit_forward(x: torch.Tensor, t100: torch.Tensor, emb: Optional[torch.Tensor], workspace: torch.Tensor, handle: int,
                latent_h: int) -> torch.Tensor:
    """Transformer.forward (model/denoiser/transformer.py:158-193) -> t2s_dit_forward."""
    _cuda(x, t100, emb, workspace)
    lib = _lib.load()
    pk = _obj(handle)
    out = torch.empty_like(x)
    B = x.shape[0]
    with torch.cuda.device(x.device):
        rc = lib.t2s_dit_forward(pk.ref, x.data_ptr(), t100.data_ptr(), _ptr(emb), out.data_ptr(), B, _aligned(workspace),
                                 lib.t2s_dit_workspace_bytes_h(B, latent_h), _stream(x.device))
    _lib.check(rc, "t2s_dit_forward")
    return out


@dit_forward.register_fake
def _(x, t100, emb, workspace, handle, latent_h):
    return torch.empty_like(x)


# ---------------------------------------------------------------------------------------------- fused guided sampling loop
@torch.library.custom_op("t2s_b200::sample_loop", mutates_args=("x", "trace", "workspace"), device_types="cuda")
def sample_loop(x: torch.Tensor, emb: torch.Tensor, t100: torch.Tensor, coef: torch.Tensor, step_noise: Optional[torch.Tensor],
                trace: Optional[torch.Tensor], workspace: torch.Tensor, kind: int, steps: int, cfg_scale: float, seed: int,
                use_seed: bool, handle: int, latent_h: int) -> None:
    """The guided loop of infer.py:76-88 on x (B,64,H) in place -> t2s_sample / t2s_sample_ddpm_seeded.
    coef: HOST float32 (steps,3) table; t100: DEVICE (steps,)."""
    _cuda(x, emb, t100, step_noise, trace, workspace)
    if coef.is_cuda or coef.dtype != torch.float32 or not coef.is_contiguous():
        raise RuntimeError("t2s_b200::sample_loop: coef must be a contiguous float32 HOST tensor (steps,3)")
    lib = _lib.load()
    pk = _obj(handle)
    B = x.shape[0]
    nbytes = lib.t2s_dit_workspace_bytes_h(2 * B, latent_h)
    coef_p = C.cast(C.c_void_p(coef.data_ptr()), C.POINTER(C.c_float))
    with torch.cuda.device(x.device):
        if use_seed:
            rc = lib.t2s_sample_ddpm_seeded(pk.ref, x.data_ptr(), emb.data_ptr(), t100.data_ptr(), coef_p, C.c_ulonglong(seed & (2 ** 64 - 1)),
                                            _ptr(trace), B, steps, float(cfg_scale), _aligned(workspace), nbytes, _stream(x.device))
            _lib.check(rc, "t2s_sample_ddpm_seeded")
        else:
            rc = lib.t2s_sample(pk.ref, kind, x.data_ptr(), emb.data_ptr(), t100.data_ptr(), coef_p, _ptr(step_noise), _ptr(trace), B, steps,
                                float(cfg_scale), _aligned(workspace), nbytes, _stream(x.device))
            _lib.check(rc, "t2s_sample")


@sample_loop.register_fake
def _(x, emb, t100, coef, step_noise, trace, workspace, kind, steps, cfg_scale, seed, use_seed, handle, latent_h):
    return None


# ---------------------------------------------------------------------------------------------- LA-VAE
@torch.library.custom_op("t2s_b200::vae_decode", mutates_args=(), device_types="cuda")
def vae_decode(z: torch.Tensor, length: int, handle: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Decoder.forward (model/pretrained/vqvae.py:97-105) before its squeeze -> t2s_vae_decode: (series (B,L), after (B,64,L/4))."""
    _cuda(z)
    pk = _obj(handle)
    B = z.shape[0]
    series = torch.empty(B, length, device=z.device, dtype=torch.float32)
    after = torch.empty(B, 64, length // 4, device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        rc = _lib.load().t2s_vae_decode(pk.ref, z.data_ptr(), series.data_ptr(), after.data_ptr(), B, int(length), _stream(z.device))
    _lib.check(rc, "t2s_vae_decode")
    return series, after


@vae_decode.register_fake
def _(z, length, handle):
    B = z.shape[0]
    return z.new_empty(B, length), z.new_empty(B, 64, length // 4)


@torch.library.custom_op("t2s_b200::vae_decode_into", mutates_args=("series",), device_types="cuda")
def vae_decode_into(z: torch.Tensor, series: torch.Tensor, handle: int) -> None:
    """t2s_vae_decode into a caller-owned (B,L) buffer, without the pre-transposed-conv activations."""
    _cuda(z, series)
    pk = _obj(handle)
    with torch.cuda.device(z.device):
        rc = _lib.load().t2s_vae_decode(pk.ref, z.data_ptr(), series.data_ptr(), None, z.shape[0], int(series.shape[1]), _stream(z.device))
    _lib.check(rc, "t2s_vae_decode")


@vae_decode_into.register_fake
def _(z, series, handle):
    return None


@torch.library.custom_op("t2s_b200::vae_encode", mutates_args=(), device_types="cuda")
def vae_encode(x: torch.Tensor, handle: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Encoder.forward (model/pretrained/vqvae.py:57-71) -> t2s_vae_encode: (z (B,64,30), before (B,64,L/4))."""
    _cuda(x)
    pk = _obj(handle)
    B, L = x.shape
    z = torch.empty(B, 64, 30, device=x.device, dtype=torch.float32)
    before = torch.empty(B, 64, L // 4, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        rc = _lib.load().t2s_vae_encode(pk.ref, x.data_ptr(), z.data_ptr(), before.data_ptr(), B, L, _stream(x.device))
    _lib.check(rc, "t2s_vae_encode")
    return z, before


@vae_encode.register_fake
def _(x, handle):
    B, L = x.shape
    return x.new_empty(B, 64, 30), x.new_empty(B, 64, L // 4)


# ---------------------------------------------------------------------------------------------- step-wise process maths
@torch.library.custom_op("t2s_b200::rf_euler", mutates_args=(), device_types="cuda")
def rf_euler(x: torch.Tensor, v: torch.Tensor, dt: float) -> torch.Tensor:
    """RectifiedFlow.euler (model/backbone/rectified_flow.py:5-7) -> t2s_rf_euler."""
    _cuda(x, v)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().t2s_rf_euler(x.data_ptr(), v.data_ptr(), float(dt), out.data_ptr(), x.numel(), _stream(x.device))
    _lib.check(rc, "t2s_rf_euler")
    return out


@rf_euler.register_fake
def _(x, v, dt):
    return torch.empty_like(x)


@torch.library.custom_op("t2s_b200::ddpm_p_sample", mutates_args=(), device_types="cuda")
def ddpm_p_sample(xt: torch.Tensor, eps_hat: torch.Tensor, noise: torch.Tensor, c1: torch.Tensor, c2: torch.Tensor,
                  c3: torch.Tensor) -> torch.Tensor:
    """DDPM.p_sample (model/backbone/DDPM.py:28-36) with the per-sample coefficients gathered by the caller -> t2s_ddpm_p_sample."""
    _cuda(xt, eps_hat, noise, c1, c2, c3)
    out = torch.empty_like(xt)
    B = xt.shape[0]
    with torch.cuda.device(xt.device):
        rc = _lib.load().t2s_ddpm_p_sample(xt.data_ptr(), eps_hat.data_ptr(), noise.data_ptr(), c1.data_ptr(), c2.data_ptr(), c3.data_ptr(),
                                           out.data_ptr(), B, xt.numel() // B, _stream(xt.device))
    _lib.check(rc, "t2s_ddpm_p_sample")
    return out


@ddpm_p_sample.register_fake
def _(xt, eps_hat, noise, c1, c2, c3):
    return torch.empty_like(xt)


@torch.library.custom_op("t2s_b200::make_inputs", mutates_args=(), device_types="cuda")
def make_inputs(kind: int, x1: torch.Tensor, noise: torch.Tensor, ca: torch.Tensor, cb: Optional[torch.Tensor],
                latent_h: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """kind 0: RectifiedFlow.create_flow (rectified_flow.py:8-12): x_t = ca x1 + (1 - ca) noise, target = x1 - noise;
    kind 1: DDPM.q_sample (DDPM.py:19-27): x_t = ca x1 + cb noise, target = noise -> t2s_train_make_inputs_h."""
    _cuda(x1, noise, ca, cb)
    xt, target = torch.empty_like(x1), torch.empty_like(x1)
    with torch.cuda.device(x1.device):
        rc = _lib.load().t2s_train_make_inputs_h(int(kind), x1.data_ptr(), noise.data_ptr(), ca.data_ptr(), _ptr(cb), xt.data_ptr(),
                                                 target.data_ptr(), x1.shape[0], int(latent_h), _stream(x1.device))
    _lib.check(rc, "t2s_train_make_inputs")
    return xt, target


@make_inputs.register_fake
def _(kind, x1, noise, ca, cb, latent_h):
    return torch.empty_like(x1), torch.empty_like(x1)


OPS = ("dit_forward", "sample_loop", "vae_decode", "vae_decode_into", "vae_encode", "rf_euler", "ddpm_p_sample", "make_inputs")
