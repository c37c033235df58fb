"""Fused classifier-free-guided sampling: the whole loop of infer.py:75-95 in one call.

``T2SSampler.sample`` enqueues, on the current CUDA stream and without any host synchronisation,
``steps`` x {conditioning, patch-embed+QKV, 4 x (attention, fused token block)} where the last block
kernel also applies the guidance mix and the Euler (rectified flow) or ancestral (DDPM) update in
place, followed by the LA-VAE decode.  Samples are independent, so multi-GPU sampling shards the
batch across ranks with no collective inside the loop and one final all_gather (``gather_series``).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .backbone import DDPM, RectifiedFlow
from .denoiser import Transformer, _aligned

KIND = {"flowmatching": 0, "rf": 0, "rectified_flow": 0, "ddpm": 1}


def _signed64(v: int) -> int:
    """64-bit pattern as the signed int a custom-op `int` argument carries."""
    return v - (1 << 64) if v >= (1 << 63) else v


class T2SSampler:
    NOISE_WINDOW_BYTES = 1 << 30          # DDPM step noise drawn per window of steps when the caller supplies none

    GRAPH_MAX_BATCH = 64                  # sample_host replays a captured CUDA graph up to this batch (launch-gap bound regime)

    def __init__(self, dit: Transformer, vae=None):
        self.dit = dit
        self.decoder = getattr(vae, "decoder", vae)
        self._tables = {}
        self._graphs = {}

    def _table(self, kind: int, steps: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (kind, steps, str(device))
        if key not in self._tables:
            if kind == 0:
                t = RectifiedFlow.timesteps(steps)
                coef = RectifiedFlow.coefficients(steps)
            else:
                t = DDPM.timesteps(steps)
                coef = DDPM.coefficients(steps)
            t100 = (t * 100.0).to(torch.float32).to(device)             # TimeEmbedding, transformer.py:31
            self._tables[key] = (t100, coef.to(torch.float32).contiguous())      # coefficients stay on the host: launch parameters
        return self._tables[key]

    @torch.no_grad()
    def sample_latent(self, emb: torch.Tensor, steps: int = 100, cfg_scale: float = 7.0, backbone: str = "flowmatching",
                      noise: Optional[torch.Tensor] = None, step_noise: Optional[torch.Tensor] = None,
                      generator: Optional[torch.Generator] = None, trace: bool = False, chunk: Optional[int] = None,
                      noise_source: str = "philox", seed: Optional[int] = None):
        """emb (B,128) CUDA -> final latent (B,64,H) [, per-step guided predictions (steps,B,64,H)].

        DDPM draws Gaussian noise inside every p_sample (DDPM.py:35).  ``step_noise`` (steps,B,64,H) supplies it explicitly
        (parity tests); otherwise ``noise_source="philox"`` (default) generates it inside the update kernel from ``seed``
        (drawn from ``generator`` / torch's global RNG when None) — one enqueue for the whole loop, no noise tensor — and
        ``noise_source="torch"`` draws it with ``torch.randn`` per window of steps."""
        if not emb.is_cuda:
            raise RuntimeError("T2SSampler needs CUDA tensors (no CPU fallback); use sample_host for host buffers")
        from . import ops
        kind = KIND[backbone]
        dev = emb.device
        B, H = emb.shape[0], self.dit.H
        emb = emb.detach().to(torch.float32).contiguous()
        if noise is None:
            x = torch.randn(B, 64, H, device=dev, dtype=torch.float32, generator=generator)    # infer.py:75
        else:
            x = noise.detach().to(device=dev, dtype=torch.float32).clone().contiguous()
            assert tuple(x.shape) == (B, 64, H), f"noise must be (B,64,{H})"
        lat = 64 * H
        if kind == 1 and step_noise is not None:
            step_noise = step_noise.detach().to(device=dev, dtype=torch.float32).contiguous()
            assert step_noise.shape == (steps, B, 64, H)
        tr = torch.empty(steps, B, 64, H, device=dev, dtype=torch.float32) if trace else None
        t100, coef = self._table(kind, steps, dev)
        pk = self.dit.packed()
        chunk = B if not chunk else min(int(chunk), B)
        ws = self.dit.workspace(2 * chunk, dev)
        # DDPM draws fresh Gaussian noise inside every p_sample (DDPM.py:35).  When the caller does not supply it, it is drawn
        # here in windows of steps (<= NOISE_WINDOW_BYTES at a time) and the loop is enqueued window by window through the
        # same C entry (t100 / coef offset by the window start): no host synchronisation, bounded memory at any batch size.
        philox = kind == 1 and step_noise is None and noise_source == "philox"
        if kind == 1 and step_noise is None and noise_source not in ("philox", "torch"):
            raise ValueError("noise_source must be 'philox' or 'torch'")
        if philox and seed is None:
            gdev = generator.device if generator is not None else "cpu"
            seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=gdev).item())
        if kind == 1 and step_noise is None and not philox:
            win = max(1, min(steps, self.NOISE_WINDOW_BYTES // (B * lat * 4)))
        else:
            win = steps
        with torch.cuda.device(dev):
            for j0 in range(0, steps, win):
                nj = min(win, steps - j0)
                sn_w = None
                if kind == 1 and not philox:
                    sn_w = step_noise[j0:j0 + nj] if step_noise is not None else \
                        torch.randn(nj, B, 64, H, device=dev, dtype=torch.float32, generator=generator)
                for b0 in range(0, B, chunk):
                    nb = min(chunk, B - b0)
                    sn = tr_c = None
                    if kind == 1 and not philox:
                        sn = sn_w[:, b0:b0 + nb].contiguous() if nb != B else sn_w
                    if trace:
                        tr_c = tr[j0:j0 + nj] if nb == B else torch.empty(nj, nb, 64, H, device=dev, dtype=torch.float32)
                    # one custom-op call (t2s_b200::sample_loop -> t2s_sample / t2s_sample_ddpm_seeded) enqueues the whole window
                    if philox:
                        # every batch chunk gets its own key so that element indices (local to a call) never repeat a stream
                        ops.sample_loop(x[b0:b0 + nb], emb[b0:b0 + nb], t100, coef, None, tr_c, ws, 1, steps, float(cfg_scale),
                                        _signed64((seed + 0x9E3779B97F4A7C15 * (b0 // chunk)) & (2 ** 64 - 1)), True, pk.handle, H)
                    else:
                        ops.sample_loop(x[b0:b0 + nb], emb[b0:b0 + nb], t100[j0:j0 + nj], coef[j0:j0 + nj], sn, tr_c, ws, kind, nj,
                                        float(cfg_scale), 0, False, pk.handle, H)
                    if trace and nb != B:
                        tr[j0:j0 + nj, b0:b0 + nb] = tr_c
        return (x, tr) if trace else x

    @torch.no_grad()
    def sample(self, emb: torch.Tensor, length: int, steps: int = 100, cfg_scale: float = 7.0,
               backbone: str = "flowmatching", noise=None, step_noise=None, generator=None, chunk=None,
               return_latent: bool = False):
        """Text embeddings (B,128) -> generated series (B,length) fp32 (infer.py:75-95)."""
        if self.decoder is None:
            raise RuntimeError("T2SSampler was built without an LA-VAE decoder")
        z = self.sample_latent(emb, steps, cfg_scale, backbone, noise, step_noise, generator, False, chunk)
        if self.dit.H == 30 and hasattr(self.decoder, "decode_into"):
            series = torch.empty(z.shape[0], int(length), device=z.device, dtype=torch.float32)
            self.decoder.decode_into(z, int(length), series, None)          # fused single-kernel T2S decoder
        else:
            # the fork's multivariate LA-VAE (mylavae.Decoder, myinfer.py:147): (B, 64, flow_dim) -> (B, input_dim, length)
            series = self.decoder(z, int(length))[0]
        return (series, z) if return_latent else series

    @torch.no_grad()
    def sample_graph(self, emb: torch.Tensor, length: int, steps: int = 100, cfg_scale: float = 7.0, noise=None, generator=None):
        """Rectified-flow `sample` through a captured CUDA graph (one graph per (batch, length, steps, cfg, weights)): the
        ~10 launches per step are replayed without per-launch gaps, which is what bounds small batches (B = 1: 15.3 -> 13.7 ms
        per 100 steps).  The loop only enqueues kernels (no allocation, no host sync), so capture needs nothing special;
        inputs go through static buffers, the result is copied out of the graph's memory."""
        if self.decoder is None or self.dit.H != 30:
            raise RuntimeError("sample_graph needs the LA-VAE decoder and the T2S shape")
        dev, B = emb.device, emb.shape[0]
        pk, dpk = self.dit.packed(), self.decoder._packed_weights()
        # a captured graph holds RAW device pointers (workspace, packed weights of both models): the cache entry keeps those
        # objects alive for as long as the graph exists, and the key uses their monotonically increasing pack generations
        # (an id() can be reused by a new object once the old pack is freed)
        key = (str(dev), B, int(length), int(steps), float(cfg_scale), pk.generation, dpk.generation)
        for k in [k for k in self._graphs if k[0] == key[0] and k[5:] != key[5:]]:
            del self._graphs[k]                                     # weights changed: graphs of the old packs can never be replayed
        ent = self._graphs.get(key)
        if ent is None:
            s_emb = torch.empty(B, 128, device=dev, dtype=torch.float32)
            s_x0 = torch.empty(B, 64, 30, device=dev, dtype=torch.float32)
            s_emb.copy_(emb)
            s_x0.normal_()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                       # warm-up outside the capture: workspace, tables, allocator
                self.sample(s_emb, length, steps=steps, cfg_scale=cfg_scale, noise=s_x0)
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s_out = self.sample(s_emb, length, steps=steps, cfg_scale=cfg_scale, noise=s_x0)
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = (g, s_emb, s_x0, s_out, (self.dit.workspace(2 * B, dev), pk, dpk))
        g, s_emb, s_x0, s_out, _keepalive = ent
        s_emb.copy_(emb.detach().to(torch.float32), non_blocking=True)
        if noise is None:
            s_x0.normal_(generator=generator)                    # infer.py:75
        else:
            s_x0.copy_(noise, non_blocking=True)
        g.replay()
        return s_out.clone()

    @torch.no_grad()
    def sample_host(self, emb_host: torch.Tensor, length: int, out_host: Optional[torch.Tensor] = None, graph="auto", **kw) -> torch.Tensor:
        """End-to-end call with HOST buffers: H2D copy of the (pinned) text embeddings, the fused loop,
        D2H copy of the series.  Returns the host tensor after synchronising the stream.  Small rectified-flow batches
        (<= GRAPH_MAX_BATCH) replay a captured CUDA graph (`graph="auto"`; pass False to force plain enqueueing)."""
        dev = next(self.dit.parameters()).device
        emb = emb_host.to(dev, non_blocking=True)
        plain = set(kw) <= {"steps", "cfg_scale", "noise", "generator", "backbone"} and KIND[kw.get("backbone", "flowmatching")] == 0
        if graph and plain and emb.shape[0] <= self.GRAPH_MAX_BATCH and self.dit.H == 30 and self.decoder is not None:
            series = self.sample_graph(emb, length, **{k: v for k, v in kw.items() if k != "backbone"})
        else:
            series = self.sample(emb, length, **kw)
        if out_host is None:
            out_host = torch.empty(series.shape, dtype=torch.float32, pin_memory=True)
        out_host.copy_(series, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_host


# ---------------------------------------------------------------------------------------- multi-GPU
def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch shard [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_series(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """The only collective of multi-GPU sampling: all_gather of the per-rank series
    (np.concatenate at infer.py:112-113).  Works with NCCL (CUDA) and gloo (CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    assert sizes[rank][1] - sizes[rank][0] == local.shape[0]
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
