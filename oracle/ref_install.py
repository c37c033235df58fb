#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the hot path into git-ignored ``baseline/_ref/``.

TEST / BASELINE INFRASTRUCTURE ONLY — nothing under t2ms_b200/ imports this file or what it stages.

The reference is pure Python, so there is nothing to compile: ``install()`` copies the five files the path needs
 byte for byte from where they lie under /root/reference, plus the two import shims the image
lacks (``timm==1.0.11`` Attention / Mlp, an empty ``matplotlib``; the same shims ``oracle/make_golden.py`` uses).
``baseline/_ref`` is listed in .gitignore (reference sources never enter the history) but not in .gpurunignore, so it
travels to the GPU box, where /root/reference does not exist.  ``bench.py --impl reference`` and the
``gpu_eager_baseline`` leg drive these modules through their own public API (``Transformer.forward``,
``RectifiedFlow.euler``, ``DDPM.p_sample``, ``vqvae.decoder``) with the loop of infer.py:75-95.

    python oracle/ref_install.py            # run in the build container; __graft_entry__.build() calls it too
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_REF = "/root/reference"
DEFAULT_DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "model/denoiser/transformer.py",
    "model/backbone/rectified_flow.py",
    "model/backbone/DDPM.py",
    "model/pretrained/core.py",
    "model/pretrained/vqvae.py",
]


def install(ref_root: str = DEFAULT_REF, dst: str = DEFAULT_DST) -> str | None:
    """Copy the reference files + shims to ``dst``; returns ``dst``, or None when the reference is not mounted."""
    if not os.path.isdir(ref_root):
        return dst if available(dst) else None
    sys.path.insert(0, ROOT)
    from oracle.make_golden import TIMM_SHIM
    for rel in FILES:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(ref_root, rel), out)
    shim = os.path.join(dst, "_shims")
    os.makedirs(os.path.join(shim, "timm", "models"), exist_ok=True)
    os.makedirs(os.path.join(shim, "matplotlib"), exist_ok=True)
    for rel in ("timm/__init__.py", "timm/models/__init__.py", "matplotlib/__init__.py", "matplotlib/pyplot.py", "matplotlib/animation.py"):
        open(os.path.join(shim, rel), "w").close()
    with open(os.path.join(shim, "timm", "models", "vision_transformer.py"), "w") as f:
        f.write(textwrap.dedent(TIMM_SHIM))
    with open(os.path.join(dst, "README"), "w") as f:
        f.write("Unmodified copies of the reference's hot-path modules (Bill9125/T2MS) staged by oracle/ref_install.py for the\n"
                "bench baselines.  Git-ignored; not product source.\n")
    return dst


def available(dst: str = DEFAULT_DST) -> bool:
    return all(os.path.exists(os.path.join(dst, rel)) for rel in FILES)


def load(dst: str = DEFAULT_DST):
    """Import the staged reference modules -> dict of the reference classes (real timm / matplotlib win over the shims)."""
    if not available(dst):
        raise RuntimeError(f"{dst} is not staged: run `python oracle/ref_install.py` where /root/reference is mounted")
    for mod in ("timm.models.vision_transformer", "matplotlib.pyplot"):
        try:
            importlib.import_module(mod)
        except Exception:
            shim = os.path.join(dst, "_shims")
            if shim not in sys.path:
                sys.path.append(shim)
    for name in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
        del sys.modules[name]                      # e.g. aliases left by t2ms_b200.compat.install()
    sys.path.insert(0, dst)
    try:
        out = {
            "Transformer": importlib.import_module("model.denoiser.transformer").Transformer,
            "RectifiedFlow": importlib.import_module("model.backbone.rectified_flow").RectifiedFlow,
            "DDPM": importlib.import_module("model.backbone.DDPM").DDPM,
            "vqvae": importlib.import_module("model.pretrained.vqvae").vqvae,
        }
    finally:
        sys.path.remove(dst)
    return out


def reference_sample(ref, dit, vae, emb, noise, steps: int, cfg_scale: float, length: int, backbone: str = "flowmatching",
                     step_noise=None):
    """The hot loop of infer.py:75-95 on the reference's own objects (no plotting, no per-step decode of batch 0).
    emb (B,128), noise (B,64,30) on the model's device -> series (B,length)."""
    import torch
    dev = emb.device
    B = emb.shape[0]
    proc = ref["RectifiedFlow"]() if backbone == "flowmatching" else ref["DDPM"](steps, dev)
    x_t = noise.clone()
    with torch.no_grad():
        for j in range(steps):
            if backbone == "flowmatching":
                t = torch.round(torch.full((B,), j * 1.0 / steps, device=dev) * steps) / steps          # infer.py:78
                pred_uncond = dit(input=x_t, t=t, text_input=None)
                pred_cond = dit(input=x_t, t=t, text_input=emb)
                pred = pred_uncond + cfg_scale * (pred_cond - pred_uncond)
                x_t = proc.euler(x_t, pred, 1.0 / steps)
            else:
                t = torch.floor(torch.full((B,), steps - 1 - j, device=dev)).long()                    # infer.py:84
                pred_uncond = dit(input=x_t, t=t, text_input=None)
                pred_cond = dit(input=x_t, t=t, text_input=emb)
                pred = pred_uncond + cfg_scale * (pred_cond - pred_uncond)
                if step_noise is None:
                    x_t = proc.p_sample(x_t, pred, t)
                else:                                   # parity runs: the same draws as the caller supplies
                    import unittest.mock as um
                    with um.patch.object(torch, "randn", lambda *a, **k: step_noise[j]):
                        x_t = proc.p_sample(x_t, pred, t)
        series, _ = vae.decoder(x_t, length=length)
    return series


def build_reference_models(ref, dit_state, vae_state, device="cpu"):
    """Reference Transformer + vqvae carrying the given state dicts (synthetic weights of t2ms_b200.synth)."""
    from argparse import Namespace
    dit = ref["Transformer"]()
    dit.load_state_dict(dit_state, strict=True)
    vae = ref["vqvae"](Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256, embedding_dim=64))
    vae.load_state_dict(vae_state, strict=True)
    return dit.to(device).eval(), vae.to(device).eval()


if __name__ == "__main__":
    d = install()
    print("staged" if d else "reference not mounted and nothing staged", d or "")
