"""Make t2ms_b200 importable under the reference's module paths.

The reference scripts do ``from model.denoiser.transformer import Transformer`` and unpickle the
LA-VAE with ``torch.load(..., weights_only=False)`` whose class path is
``model.pretrained.vqvae.vqvae`` (infer.py:39-41, train.py:22).  ``install()`` registers alias
modules so both keep working with the B200 implementations; ``convert_reference_vae`` turns an
unpickled module / state dict into a t2ms_b200 ``vqvae``.
"""
from __future__ import annotations

import sys
import types
from argparse import Namespace

_ALIASES = {
    "model.denoiser.transformer": ("t2ms_b200.denoiser", ["Transformer", "Transformerlayer", "TimeEmbedding", "modulate",
                                                          "get_sinusoidal_positional_embeddings", "InverseLatentEmbedding"]),
    "model.backbone.rectified_flow": ("t2ms_b200.backbone", ["RectifiedFlow"]),
    "model.backbone.DDPM": ("t2ms_b200.backbone", ["DDPM", "gather"]),
    "model.pretrained.core": ("t2ms_b200.lavae", ["BaseModel"]),
    "model.pretrained.vqvae": ("t2ms_b200.lavae", ["vqvae", "Encoder", "Decoder", "Residual", "ResidualStack"]),
    # the fork's variants (mytrain.py:8-9, myinfer.py:14-15, pretrained_mylavae.py:5)
    "model.denoiser.mytransformer": ("t2ms_b200.denoiser", ["Transformer", "Transformerlayer", "TimeEmbedding", "modulate",
                                                            "get_sinusoidal_positional_embeddings", "InverseLatentEmbedding"]),
    "model.pretrained.myvqvae": ("t2ms_b200.mylavae", ["vqvae", "Encoder", "Decoder"]),
}


def install(force: bool = False) -> None:
    """Alias the B200 implementations under the reference's module paths.  When the reference's own ``model`` package is
    importable (an untouched infer.py / train.py run from the reference checkout), the REAL packages stay in place and only
    the hot-path submodules are overridden in ``sys.modules`` — ``from model.denoiser.mlp import MLP`` (infer.py:4, train.py:9)
    and every other real submodule keep importing.  Otherwise empty namespace packages stand in."""
    import importlib
    import importlib.util
    for pkg in ("model", "model.denoiser", "model.backbone", "model.pretrained"):
        if pkg in sys.modules and not force:
            continue
        real = None
        try:
            if importlib.util.find_spec(pkg) is not None:
                real = importlib.import_module(pkg)
        except (ImportError, ValueError):
            real = None
        if real is None:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
            parent, _, leaf = pkg.rpartition(".")
            if parent:
                setattr(sys.modules[parent], leaf, m)
    for alias, (target, names) in _ALIASES.items():
        if alias in sys.modules and not force:
            continue
        src = importlib.import_module(target)
        m = types.ModuleType(alias)
        for n in names:
            setattr(m, n, getattr(src, n))
        sys.modules[alias] = m
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, m)


VAE_ARGS = Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256, embedding_dim=64)


def convert_reference_vae(obj):
    """nn.Module (e.g. the unpickled reference vqvae) or state dict -> t2ms_b200.lavae.vqvae."""
    from .lavae import vqvae
    sd = obj.state_dict() if hasattr(obj, "state_dict") else obj
    m = vqvae(VAE_ARGS)
    m.load_state_dict(sd, strict=True)
    return m.eval()
