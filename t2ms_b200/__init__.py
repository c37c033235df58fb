"""t2ms_b200 — B200-native (sm_100a) implementation of the T2S generation hot path.

Public surface mirrors the reference modules used by infer.py / train.py:
``Transformer`` (model/denoiser/transformer.py), ``RectifiedFlow`` / ``DDPM`` (model/backbone),
``vqvae`` / ``Encoder`` / ``Decoder`` (model/pretrained/vqvae.py), plus ``T2SSampler`` which runs the
whole guided sampling loop + decode in one enqueue.
"""
from .backbone import DDPM, RectifiedFlow
from .denoiser import Transformer, Transformerlayer, TimeEmbedding
from .lavae import Decoder, Encoder, vqvae
from .sampler import T2SSampler, gather_series, shard_range
from .training import DitTrainer
from .outputs import load_generation, run_inference, save_generation, series_metrics

__all__ = ["Transformer", "Transformerlayer", "TimeEmbedding", "RectifiedFlow", "DDPM", "vqvae", "Encoder", "Decoder",
           "T2SSampler", "gather_series", "shard_range", "DitTrainer",
           "series_metrics", "run_inference", "save_generation", "load_generation"]
__version__ = "0.1.0"
from . import ops  # noqa: E402,F401  registers the torch.library custom ops (namespace t2s_b200)
