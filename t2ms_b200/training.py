"""Training step of the T2S-DiT (train.py:66-87) on the sm_100a kernels.

Two ways in, both ending in the same C-ABI calls (include/t2s_b200.h, "Training step"):

* drop-in: ``Transformer.forward`` in training mode returns a tensor with a ``grad_fn``
  (``dit_forward_autograd``), so the reference loop ``pred = model(x_t, t, emb); loss = mse(pred, target);
  loss.backward(); optimizer.step()`` (train.py:83-87) runs unchanged; the forward is
  ``t2s_dit_train_forward`` and ``loss.backward()`` lands in ``t2s_dit_train_backward``;
* fused: ``DitTrainer.step`` enqueues create_flow / q_sample, forward, MSE, backward, the data-parallel
  gradient all-reduce (NCCL, one flat 3.7 MB bucket) and AdamW without any host synchronisation.

There is no PyTorch fallback: the backward is the hand-written kernel chain, CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

N_LAYER = 4
_TOP = {"conv.weight": "conv_w", "conv.bias": "conv_b", "patch_emb.weight": "pe_w", "patch_emb.bias": "pe_b",
        "ln.weight": "ln_w", "ln.bias": "ln_b", "linear_emb_to_patch.weight": "lf_w", "linear_emb_to_patch.bias": "lf_b"}
_LAYER = {"attn.qkv.weight": "qkv_w", "attn.qkv.bias": "qkv_b", "attn.proj.weight": "proj_w", "attn.proj.bias": "proj_b",
          "mlp.fc1.weight": "fc1_w", "mlp.fc1.bias": "fc1_b", "mlp.fc2.weight": "fc2_w", "mlp.fc2.bias": "fc2_b",
          "adaLN_modulation.1.weight": "ada_w", "adaLN_modulation.1.bias": "ada_b"}


def trainable_names() -> List[str]:
    """The 48 parameters that receive a gradient: everything except the fixed ``pos_embed``, the
    unused ``unpatch.*`` head (transformer.py:150) and the attached frozen ``encoder.*`` (train.py:31-33)."""
    names = list(_TOP)
    for l in range(N_LAYER):
        names += [f"layers.{l}.{k}" for k in _LAYER]
    return names


def _struct(tensors: Dict[str, torch.Tensor], pos: Optional[torch.Tensor], freqs: Optional[torch.Tensor], latent_h: int = 30) -> _lib.DitParams:
    st = _lib.DitParams()
    st.latent_h = int(latent_h)
    for name, field in _TOP.items():
        setattr(st, field, tensors[name].data_ptr())
    for l in range(N_LAYER):
        for k, field in _LAYER.items():
            getattr(st, field)[l] = tensors[f"layers.{l}.{k}"].data_ptr()
    st.pos = pos.data_ptr() if pos is not None else None
    st.freqs = freqs.data_ptr() if freqs is not None else None
    return st


_FREQS: Dict[str, torch.Tensor] = {}


def _freqs(device) -> torch.Tensor:
    key = str(device)
    if key not in _FREQS:
        _FREQS[key] = torch.pow(10000, torch.linspace(0, 1, 64)).to(device=device, dtype=torch.float32).contiguous()   # transformer.py:34
    return _FREQS[key]


def _aligned_ptr(buf: torch.Tensor) -> int:
    return (buf.data_ptr() + 255) & ~255


class _Workspace:
    def __init__(self):
        self.buf = None
        self.nseq = 0
        self.generation = 0          # bumped by every training-mode forward: backward() checks it still owns the activations

    def get(self, nseq: int, device, latent_h: int = 30) -> Tuple[int, int]:
        lib = _lib.load()
        nbytes = lib.t2s_train_workspace_bytes_h(nseq, latent_h)
        if nbytes == 0:
            raise RuntimeError("t2s_train_workspace_bytes_h: " + lib.t2s_last_error().decode(errors="replace"))
        if self.buf is None or self.buf.numel() < nbytes + 256 or self.buf.device != device:
            self.buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            self.nseq = nseq
        return _aligned_ptr(self.buf), nbytes


_MODULE_WS = weakref.WeakKeyDictionary()


def _module_ws(model) -> _Workspace:
    if model not in _MODULE_WS:
        _MODULE_WS[model] = _Workspace()
    return _MODULE_WS[model]


class FlatBuffer:
    """One contiguous fp32 buffer holding every trainable tensor at a 256-byte aligned offset."""

    def __init__(self, shapes: Dict[str, torch.Size], device, tail: int = 0):
        self.offsets, off = {}, 0
        for n, s in shapes.items():
            self.offsets[n] = (off, s)
            off += (s.numel() + 63) // 64 * 64
        self.numel = off                                  # elements holding tensors; `tail` extra fp32 slots follow
        self.flat = torch.zeros(off + tail, dtype=torch.float32, device=device)

    def view(self, name: str) -> torch.Tensor:
        off, s = self.offsets[name]
        return self.flat[off:off + s.numel()].view(s)

    def views(self) -> Dict[str, torch.Tensor]:
        return {n: self.view(n) for n in self.offsets}


def _own_trainable(model) -> Dict[str, torch.nn.Parameter]:
    named = dict(model.named_parameters())
    return {n: named[n] for n in trainable_names()}


def _prep_inputs(x, t, text, latent_h: int = 30):
    if not x.is_cuda:
        raise RuntimeError("t2ms_b200 training needs CUDA tensors (no CPU / PyTorch fallback)")
    B = x.shape[0]
    assert tuple(x.shape[1:]) == (64, latent_h), f"latent must be (B,64,{latent_h}), got {tuple(x.shape)}"
    x = x.detach().to(torch.float32).contiguous()
    t100 = (t.detach() * 100.0).to(device=x.device, dtype=torch.float32).contiguous()       # transformer.py:31
    assert t100.shape == (B,)
    if text is not None:
        text = text.detach().to(device=x.device, dtype=torch.float32).contiguous()
        assert tuple(text.shape) == (B, 128)
    return x, t100, text


class _DitFunction(torch.autograd.Function):
    """pred = Transformer.forward(x_t, t, text) with the hand-written backward."""

    @staticmethod
    def forward(ctx, model, x, t100, text, *params):
        lib = _lib.load()
        names = trainable_names()
        tensors = {n: p.detach() for n, p in zip(names, params)}
        for n, p in tensors.items():
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError(f"parameter {n} must be a contiguous fp32 CUDA tensor")
        dev = x.device
        st = _struct(tensors, model.pos_embed.detach(), _freqs(dev), model.H)
        ws = _module_ws(model)
        ptr, nbytes = ws.get(x.shape[0], dev, model.H)
        pred = torch.empty_like(x)
        with torch.cuda.device(dev):
            rc = lib.t2s_dit_train_forward(C.byref(st), x.data_ptr(), t100.data_ptr(), text.data_ptr() if text is not None else None,
                                           pred.data_ptr(), x.shape[0], ptr, nbytes, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_dit_train_forward")
        ws.generation += 1
        ctx.model, ctx.st, ctx.nseq, ctx.keep = model, st, x.shape[0], (tensors, ws.buf)
        ctx.generation = ws.generation
        ctx.versions = tuple(p._version for p in params)
        ctx.params = params
        ctx.shapes = {n: p.shape for n, p in tensors.items()}
        return pred

    @staticmethod
    def backward(ctx, dpred):
        lib = _lib.load()
        model = ctx.model
        ws = _module_ws(model)
        if ctx.keep[1] is not ws.buf or ctx.generation != ws.generation:
            raise RuntimeError("t2ms_b200.Transformer keeps the activations of ONE training-mode forward per module: another "
                               "forward ran before this backward() (e.g. (loss1 + loss2).backward() over two forwards, or gradient "
                               "accumulation without a backward in between).  Call backward() after each forward, or use "
                               "DitTrainer(micro_batch=...) for accumulation.")
        if tuple(p._version for p in ctx.params) != ctx.versions:
            raise RuntimeError("a parameter of t2ms_b200.Transformer was modified in place between forward and backward()")
        dev = dpred.device
        gbuf = FlatBuffer(ctx.shapes, dev)
        gst = _struct(gbuf.views(), None, None, model.H)
        dpred = dpred.to(torch.float32).contiguous()
        ptr, nbytes = _module_ws(model).get(ctx.nseq, dev, model.H)
        with torch.cuda.device(dev):
            rc = lib.t2s_dit_train_backward(C.byref(ctx.st), C.byref(gst), dpred.data_ptr(), ctx.nseq, ptr, nbytes,
                                            torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_dit_train_backward")
        return (None, None, None, None, *[gbuf.view(n) for n in trainable_names()])


def dit_forward_autograd(model, x, t, text):
    """Transformer.forward in training mode (called from t2ms_b200.denoiser.Transformer.forward)."""
    x, t100, text = _prep_inputs(x, t, text, model.H)
    params = [p for p in _own_trainable(model).values()]
    return _DitFunction.apply(model, x, t100, text, *params)


class DitTrainer:
    """Fused training step: forward + MSE + backward + (data-parallel all-reduce) + AdamW, all enqueued on the
    current stream.  Mirrors train.py:37 (AdamW lr 1e-4, weight_decay 0) and train.py:66-87.

    The trainable parameters are re-homed into one flat fp32 buffer (the module's ``nn.Parameter`` objects become
    views of it, state-dict names and shapes unchanged) so that the optimizer and the gradient all-reduce are one
    kernel / one collective over 3.7 MB.
    """

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, group=None,
                 coin_seed: int = 0):
        params = _own_trainable(model)
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("DitTrainer needs the model on a CUDA device (no CPU fallback)")
        self.model, self.device, self.group = model, dev, group
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        shapes = {n: p.shape for n, p in params.items()}
        # the gradient bucket carries the loss sum in its tail: the data-parallel exchange is ONE all_reduce
        self.params, self.grads = FlatBuffer(shapes, dev), FlatBuffer(shapes, dev, tail=64)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.params.flat), torch.zeros_like(self.params.flat)
        with torch.no_grad():
            for n, p in params.items():
                v = self.params.view(n)
                v.copy_(p.data)
                p.data = v
                p.grad = self.grads.view(n)
        self.H = int(getattr(model, "H", 30))
        self.lat = 64 * self.H
        self._pst = _struct(self.params.views(), model.pos_embed.detach(), _freqs(dev), self.H)
        self._gst = _struct(self.grads.views(), None, None, self.H)
        self.ws = _Workspace()
        self.loss_sum = self.grads.flat[self.grads.numel:self.grads.numel + 1]       # [sum of squared errors]
        self.step_count = 0
        self._graphs = {}
        # the per-batch CFG-dropout coin (train.py:80, CPU RNG): with data parallelism every rank draws it from an identically
        # seeded CPU generator — no broadcast, no device round trip; a single process keeps the reference's global-RNG draw
        self._coin_gen = torch.Generator().manual_seed(int(coin_seed))

    # ------------------------------------------------------------------ pieces
    def zero_grad(self):
        self.grads.flat.zero_()                           # gradients + loss sum + element count (the bucket's tail)

    def forward_backward(self, x_t, t, emb, target, loss_numel: Optional[float] = None, backward: bool = True, pred=None):
        """Accumulates dL/dparam into ``self.grads`` and sum((pred-target)^2) into ``self.loss_sum``."""
        x_t, t100, emb = _prep_inputs(x_t, t, emb, self.H)
        target = target.detach().to(torch.float32).contiguous()
        self._fb_raw(x_t, t100, emb, target, loss_numel, backward, pred)

    def _fb_raw(self, x_t, t100, emb, target, loss_numel=None, backward: bool = True, pred=None):
        """forward_backward on prepared (fp32, contiguous, t already x100) tensors: one C call, nothing else enqueued."""
        lib = _lib.load()
        B = x_t.shape[0]
        numel = float(loss_numel if loss_numel is not None else B * self.lat)
        ptr, nbytes = self.ws.get(B, self.device, self.H)
        with torch.cuda.device(self.device):
            rc = lib.t2s_dit_train_step(C.byref(self._pst), C.byref(self._gst) if backward else None, x_t.data_ptr(), t100.data_ptr(),
                                        emb.data_ptr() if emb is not None else None, target.data_ptr(), self.loss_sum.data_ptr(),
                                        pred.data_ptr() if pred is not None else None, B, numel, ptr, nbytes,
                                        torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_dit_train_step")

    # ------------------------------------------------------------------ the ~125 launches of a step as one CUDA graph
    GRAPHS = True                      # class-wide switch (tests compare both paths)

    def _zero_forward_backward(self, x_t, t, emb, target, numel: float):
        """zero_grad + forward + backward of one whole (non-accumulated) batch.  The C entry only enqueues kernels into a
        caller-owned workspace, so from the second call with the same (batch, text | no text, normalisation) on it is replayed
        from a CUDA graph over static input buffers: the step is ~130 short kernels and was launch-gap bound (0.5 of 6.8 ms)."""
        B = x_t.shape[0]
        if not self.GRAPHS or torch.cuda.is_current_stream_capturing():
            self.zero_grad()
            self.forward_backward(x_t, t, emb, target, loss_numel=numel)
            return
        x_t, t100, emb = _prep_inputs(x_t, t, emb, self.H)
        target = target.detach().to(torch.float32).contiguous()
        ptr, _ = self.ws.get(B, self.device, self.H)
        key = (B, emb is not None, float(numel), ptr)
        ent = self._graphs.get(key)
        if ent is None:
            # first call: run eagerly (this IS the step), then record the same enqueue sequence for the next calls
            self.zero_grad()
            self._fb_raw(x_t, t100, emb, target, numel)
            st = dict(x=torch.empty_like(x_t), t=torch.empty_like(t100), e=torch.empty_like(emb) if emb is not None else None,
                      y=torch.empty_like(target))
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g):
                self.zero_grad()
                self._fb_raw(st["x"], st["t"], st["e"], st["y"], numel)
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = (g, st, self.ws.buf)
            return
        g, st, _ws = ent
        st["x"].copy_(x_t, non_blocking=True); st["t"].copy_(t100, non_blocking=True); st["y"].copy_(target, non_blocking=True)
        if emb is not None:
            st["e"].copy_(emb, non_blocking=True)
        g.replay()

    def allreduce_grads(self):
        """Data-parallel exchange: ONE SUM all-reduce of the flat gradient bucket, whose tail carries the loss sum."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.SUM, group=self.group)

    def optimizer_step(self, lr: Optional[float] = None, grad_scale: float = 1.0):
        lib = _lib.load()
        self.step_count += 1
        with torch.cuda.device(self.device):
            rc = lib.t2s_adamw_step(self.params.flat.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                                    self.exp_avg_sq.data_ptr(), self.params.flat.numel(), self.step_count,
                                    float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    float(grad_scale), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_adamw_step")
        self.model._packed = None                     # inference weight images are stale now (packed() re-packs, new generation)

    # ------------------------------------------------------------------ inputs (train.py:68-76)
    def make_inputs(self, backbone: str, x1, noise, t, ddpm=None):
        """RF: x_t = t x1 + (1-t) x0, target = x1 - x0 (rectified_flow.py:8-12, train.py:71);
        DDPM: x_t = sqrt(ab_t) x1 + sqrt(1-ab_t) eps, target = eps (DDPM.py:19-27)."""
        lib = _lib.load()
        x1 = x1.detach().to(torch.float32).contiguous()
        noise = noise.detach().to(torch.float32).contiguous()
        xt, target = torch.empty_like(x1), torch.empty_like(x1)
        if backbone in ("flowmatching", "rf", "rectified_flow"):
            kind, ca, cb = 0, t.detach().to(device=x1.device, dtype=torch.float32).contiguous(), None
        else:
            ab = ddpm.alpha_bar.to(x1.device).gather(-1, t.to(x1.device))
            kind, ca, cb = 1, (ab ** 0.5).contiguous(), ((1 - ab) ** 0.5).contiguous()
        with torch.cuda.device(x1.device):
            rc = lib.t2s_train_make_inputs_h(kind, x1.data_ptr(), noise.data_ptr(), ca.data_ptr(), cb.data_ptr() if cb is not None else None,
                                             xt.data_ptr(), target.data_ptr(), x1.shape[0], self.H, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_train_make_inputs")
        return xt, target

    # ------------------------------------------------------------------ one optimizer step
    def step(self, x_t, t, emb, target, lr: Optional[float] = None, micro_batch: Optional[int] = None,
             global_batch: Optional[int] = None) -> torch.Tensor:
        """One optimizer step on (x_t, t, emb | None, target); returns the (global) loss as a device scalar."""
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        B = x_t.shape[0]
        # equal shards (the data-parallel contract of bench.py / train_loop.py): the global element count is known on the host.
        # `global_batch` overrides it for uneven shards (every rank passes the same total).
        numel = float((global_batch if global_batch else B * world) * self.lat)
        mb = B if not micro_batch else min(int(micro_batch), B)
        if mb >= B:
            self._zero_forward_backward(x_t, t, emb, target, numel)
        else:
            self.zero_grad()
            for b0 in range(0, B, mb):
                sl = slice(b0, min(B, b0 + mb))
                self.forward_backward(x_t[sl], t[sl], emb[sl] if emb is not None else None, target[sl], loss_numel=numel)
        self.allreduce_grads()
        self.optimizer_step(lr)
        return self.loss_sum[0] / numel

    # ------------------------------------------------------------------ train.py:60-87 for one (sub-)batch
    def train_batch(self, series, emb, backbone: str = "flowmatching", total_step: int = 100, p_uncond: float = 0.3,
                    encoder=None, ddpm=None, lr: Optional[float] = None, generator: Optional[torch.Generator] = None,
                    micro_batch: Optional[int] = None, global_batch: Optional[int] = None) -> torch.Tensor:
        """One optimizer step exactly as the reference loop does it for a length-grouped sub-batch:
        frozen LA-VAE encoder (train.py:66), t / noise draws and create_flow | q_sample (:68-76), the per-BATCH
        classifier-free-guidance dropout coin (:80-82, shared by all data-parallel ranks), forward, MSE, backward,
        gradient all-reduce, AdamW (:83-87).  ``series`` (B,L) [or (B,input_dim,L) with the fork's multivariate encoder,
        mytrain.py] or an already encoded latent (B,64,H)."""
        import torch.distributed as dist
        dev = self.device
        enc = encoder if encoder is not None else getattr(self.model, "encoder", None)
        with torch.no_grad():
            is_latent = series.dim() == 3 and tuple(series.shape[1:]) == (64, self.H)
            x1 = series if is_latent else enc(series.to(dev))[0]
        B = x1.shape[0]
        if backbone in ("flowmatching", "rf", "rectified_flow"):
            t = torch.round(torch.rand(B, device=dev, generator=generator) * total_step) / total_step          # train.py:69
        else:
            t = torch.floor(torch.rand(B, device=dev, generator=generator) * total_step).long()               # train.py:73
        noise = torch.randn(x1.shape, device=dev, generator=generator)
        x_t, target = self.make_inputs(backbone, x1, noise, t, ddpm)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            coin = torch.rand(1, generator=self._coin_gen)          # same stream on every rank: no collective, no host sync
        else:
            coin = torch.rand(1)                                    # CPU RNG, train.py:80
        text = None if coin.item() < p_uncond else emb
        return self.step(x_t, t.to(torch.float32) if t.dtype != torch.float32 else t, text, target, lr=lr, micro_batch=micro_batch,
                         global_batch=global_batch)
