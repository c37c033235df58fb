"""The step after the hot path: the reference's generation file formats (infer.py:100-123) and the on-device
evaluation of generated series (evaluation.py:166-206 MSE / WAPE, driven at evaluation.py:292-300).

``run_inference`` is the loop of infer.py:66-123 on the fused sampler: per test batch encode (latent kept for
``x_t_latent_enc_array``), sample with classifier-free guidance, decode; the per-batch arrays go through the same
``squeeze`` + ``np.concatenate`` + ``[:, :, np.newaxis]`` as the reference, so the saved ``.npy`` files have the
reference's shapes: ``x_1.npy`` / ``x_t.npy`` (N, L, 1), ``x_t_latent_dec_array.npy`` / ``x_t_latent_enc_array.npy``
(N, 64, 30).  Plots / GIF frames (infer.py:90-99,157-198) are not produced.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

FILES = ("x_1.npy", "x_t.npy", "x_t_latent_dec_array.npy", "x_t_latent_enc_array.npy")


def series_metrics(ori: torch.Tensor, gen: torch.Tensor, return_per_sample: bool = False):
    """MSE and WAPE of evaluation.py:166-206 for univariate series, computed on the device.

    ori, gen: CUDA tensors (N, L) or (N, L, 1) (the saved layout) or (N, 1, L) (the layout evaluation.py:295-296
    transposes to).  Returns {"MSE", "WAPE", "valid"} as Python floats (one 24-byte D2H read)."""
    if not (ori.is_cuda and gen.is_cuda):
        raise RuntimeError("series_metrics needs CUDA tensors (no CPU fallback)")

    def two_d(x):
        if x.dim() == 3:
            if 1 not in (x.shape[1], x.shape[2]):
                raise ValueError("series_metrics handles univariate series: (N, L), (N, L, 1) or (N, 1, L)")
            x = x.reshape(x.shape[0], -1)
        return x.detach().to(torch.float32).contiguous()
    a, b = two_d(ori), two_d(gen)
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError(f"shape mismatch {tuple(ori.shape)} vs {tuple(gen.shape)}")
    n, L = a.shape
    per = torch.empty(n, 3, device=a.device, dtype=torch.float32)
    out = torch.empty(3, device=a.device, dtype=torch.float64)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.t2s_series_metrics(a.data_ptr(), b.data_ptr(), n, L, per.data_ptr(), out.data_ptr(),
                                    torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "t2s_series_metrics")
    mse, wape, valid = out.tolist()
    res = {"MSE": mse, "WAPE": wape, "valid": int(valid)}
    return (res, per) if return_per_sample else res


def stack_generation(x_1_list: Sequence[np.ndarray], x_t_list: Sequence[np.ndarray], dec_list: Sequence[np.ndarray],
                     enc_list: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """infer.py:100-116: per-batch arrays (already ``squeeze``-d like infer.py:100-108) -> the four saved arrays."""
    x_1 = np.concatenate(list(x_1_list), axis=0)[:, :, np.newaxis]
    x_t = np.concatenate(list(x_t_list), axis=0)[:, :, np.newaxis]
    return x_1, x_t, np.concatenate(list(dec_list), axis=0), np.concatenate(list(enc_list), axis=0)


def save_generation(path: str, x_1: np.ndarray, x_t: np.ndarray, latent_dec: np.ndarray, latent_enc: np.ndarray) -> None:
    """infer.py:117-121: the four ``.npy`` files of one run directory."""
    os.makedirs(path, exist_ok=True)
    for name, arr in zip(FILES, (x_1, x_t, latent_dec, latent_enc)):
        np.save(os.path.join(path, name), arr)


def load_generation(path: str) -> Dict[str, np.ndarray]:
    return {name[:-4]: np.load(os.path.join(path, name)) for name in FILES}


@torch.no_grad()
def run_inference(sampler, vae, batches: Iterable, backbone: str = "flowmatching", total_step: int = 100, cfg_scale: float = 7.0,
                  save_path: Optional[str] = None, generator: Optional[torch.Generator] = None, evaluate: bool = True):
    """The generation loop of infer.py:66-123 over an iterable of ``(text, x_1 (B, L), embedding (B, 128))`` batches
    (the dataloader's items, datafactory/dataloader.py).  Returns ``(x_1, x_t, x_t_latent_dec_array,
    x_t_latent_enc_array, metrics)``; ``metrics`` = on-device MSE / WAPE of the generated against the ground-truth
    series (evaluation.py:298) or None."""
    dev = next(sampler.dit.parameters()).device
    x1s, xts, decs, encs = [], [], [], []
    sums = None
    for _, x_1, emb in batches:
        x_1 = torch.as_tensor(x_1).float().to(dev)
        emb = torch.as_tensor(emb).float().to(dev)
        z_enc, _ = vae.encoder(x_1)                                           # infer.py:73-74
        series, z = sampler.sample(emb, x_1.shape[-1], steps=total_step, cfg_scale=cfg_scale, backbone=backbone,
                                   generator=generator, return_latent=True)   # infer.py:75-95
        if evaluate:
            _, per = series_metrics(x_1, series, return_per_sample=True)
            sums = per if sums is None else torch.cat([sums, per], 0)
        x1s.append(x_1.cpu().numpy().squeeze())                               # infer.py:100-108 (squeeze quirk kept)
        xts.append(series.cpu().numpy().squeeze())
        decs.append(z.cpu().numpy().squeeze())
        encs.append(z_enc.cpu().numpy().squeeze())
    x_1, x_t, dec, enc = stack_generation(x1s, xts, decs, encs)
    if save_path is not None:
        save_generation(save_path, x_1, x_t, dec, enc)
    metrics = None
    if evaluate and sums is not None:
        L = x_1.shape[1]
        s = sums.double()
        den = s[:, 2]
        ok = den != 0
        metrics = {"MSE": float((s[:, 0] / L).mean()), "WAPE": float((s[ok, 1] / den[ok]).mean()) if bool(ok.any()) else float("nan"),
                   "valid": int(ok.sum())}
    return x_1, x_t, dec, enc, metrics
