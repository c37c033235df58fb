"""Host side of the generic LA-VAE path (include/t2s_b200.h: t2s_lavae_*): the training step of
``vqvae.shared_eval(batch, optimizer, 'train')`` (model/pretrained/vqvae.py:118-135, myvqvae.py:116-136) and the
generic encoder / decoder forwards the fork's multivariate ``myvqvae`` modules use.

Parameters stay ordinary ``nn.Parameter`` tensors in the reference layouts; the C ABI reads them in place and
accumulates gradients straight into ``param.grad`` (allocated zero when absent), so the caller's own optimizer
(``optimizer.step()``, vqvae.py:128) sees exactly what ``loss.backward()`` would have left.  CUDA only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_FIELDS = {
    "enc_conv1_w": "encoder._conv_1.weight", "enc_conv1_b": "encoder._conv_1.bias",
    "enc_conv2_w": "encoder._conv_2.weight", "enc_conv2_b": "encoder._conv_2.bias",
    "enc_conv3_w": "encoder._conv_3.weight", "enc_conv3_b": "encoder._conv_3.bias",
    "enc_pre_w": "encoder._pre_vq_conv.weight", "enc_pre_b": "encoder._pre_vq_conv.bias",
    "dec_conv1_w": "decoder._conv_1.weight", "dec_conv1_b": "decoder._conv_1.bias",
    "dec_ct1_w": "decoder._conv_trans_1.weight", "dec_ct1_b": "decoder._conv_trans_1.bias",
    "dec_ct2_w": "decoder._conv_trans_2.weight", "dec_ct2_b": "decoder._conv_trans_2.bias",
}


def arch_of(params: Dict[str, torch.Tensor], flow_dim: int) -> Dict[str, int]:
    """Architecture fields of t2s_lavae_params read off the parameter shapes."""
    w1 = params["encoder._conv_1.weight"]
    n_res = sum(1 for k in params if k.startswith("encoder._residual_stack._layers.") and k.endswith("_block.1.weight"))
    return {"in_channels": w1.shape[1], "hidden": w1.shape[0] * 2,
            "res_hidden": params["encoder._residual_stack._layers.0._block.1.weight"].shape[0] if n_res else 1,
            "emb": params["encoder._pre_vq_conv.weight"].shape[0], "n_res": n_res, "flow_dim": int(flow_dim)}


def make_struct(tensors: Dict[str, torch.Tensor], arch: Dict[str, int]) -> _lib.LavaeParams:
    st = _lib.LavaeParams()
    for k, v in arch.items():
        setattr(st, k, int(v))
    for f, name in _FIELDS.items():
        t = tensors.get(name)
        if t is not None:
            setattr(st, f, t.data_ptr())
    for side in ("enc", "dec"):
        mod = "encoder" if side == "enc" else "decoder"
        for i in range(arch["n_res"]):
            for f, blk in (("res_w3", 1), ("res_w1", 3)):
                t = tensors.get(f"{mod}._residual_stack._layers.{i}._block.{blk}.weight")
                if t is not None:
                    getattr(st, f"{side}_{f}")[i] = t.data_ptr()
    return st


class LavaeEngine:
    """Workspace cache + struct builder of one LA-VAE module (``encoder.*`` / ``decoder.*`` parameters)."""

    def __init__(self, module, flow_dim: int):
        self.module = module
        self.flow_dim = int(flow_dim)
        self._ws: Dict[Tuple, torch.Tensor] = {}

    def _params(self, part: Optional[str] = None) -> Dict[str, torch.nn.Parameter]:
        ps = {n: p for n, p in self.module.named_parameters() if n.startswith(("encoder.", "decoder."))}
        dev = next(iter(ps.values())).device
        if dev.type != "cuda":
            raise RuntimeError("t2ms_b200 LA-VAE runs on CUDA (sm_100a) only (no CPU fallback)")
        for n, p in ps.items():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError(f"LA-VAE parameter {n} must be contiguous fp32")
        return ps

    def _workspace(self, st, B: int, L: int, dev) -> Tuple[int, int, torch.Tensor]:
        lib = _lib.load()
        nbytes = lib.t2s_lavae_workspace_bytes(C.byref(st), B, L)
        if nbytes == 0:
            raise RuntimeError("t2s_lavae_workspace_bytes: " + lib.t2s_last_error().decode(errors="replace"))
        key = (str(dev), B, L)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes + 256:
            if len(self._ws) > 4:
                self._ws.clear()
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        return (ws.data_ptr() + 255) & ~255, nbytes, ws

    def encode(self, x: torch.Tensor):
        """x (B, C, L) -> z (B, E, flow_dim), before (B, E, n)."""
        if not x.is_cuda:
            raise RuntimeError("t2ms_b200 LA-VAE needs CUDA tensors (no CPU fallback)")
        ps = self._params()
        arch = arch_of(ps, self.flow_dim)
        st = make_struct(ps, arch)
        B, L = x.shape[0], x.shape[-1]
        x = x.detach().to(torch.float32).reshape(B, arch["in_channels"], L).contiguous()
        n = ((L + 2 - 4) // 2 + 1 + 2 - 4) // 2 + 1
        z = torch.empty(B, arch["emb"], self.flow_dim, device=x.device, dtype=torch.float32)
        before = torch.empty(B, arch["emb"], n, device=x.device, dtype=torch.float32)
        wp, nb, _ = self._workspace(st, B, L, x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().t2s_lavae_encode(C.byref(st), x.data_ptr(), z.data_ptr(), before.data_ptr(), B, L, wp, nb,
                                              torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_lavae_encode")
        return z, before

    def decode(self, z: torch.Tensor, length: int):
        """z (B, E, flow_dim) -> recon (B, C, length), after (B, E, length // 4)."""
        if not z.is_cuda:
            raise RuntimeError("t2ms_b200 LA-VAE needs CUDA tensors (no CPU fallback)")
        ps = self._params()
        arch = arch_of(ps, self.flow_dim)
        st = make_struct(ps, arch)
        B, L = z.shape[0], int(length)
        z = z.detach().to(torch.float32).contiguous()
        assert tuple(z.shape) == (B, arch["emb"], self.flow_dim), f"latent must be (B,{arch['emb']},{self.flow_dim})"
        recon = torch.empty(B, arch["in_channels"], L, device=z.device, dtype=torch.float32)
        after = torch.empty(B, arch["emb"], L // 4, device=z.device, dtype=torch.float32)
        wp, nb, _ = self._workspace(st, B, L, z.device)
        with torch.cuda.device(z.device):
            rc = _lib.load().t2s_lavae_decode(C.byref(st), z.data_ptr(), recon.data_ptr(), after.data_ptr(), B, L, wp, nb,
                                              torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_lavae_decode")
        return recon, after

    def step(self, batch: torch.Tensor, backward: bool):
        """Forward (+ backward into ``param.grad``) of loss = mse(recon, batch) + mse(before, after).
        Returns (loss, recon_error, recon (B, C, L), z) as device tensors."""
        ps = self._params()
        arch = arch_of(ps, self.flow_dim)
        st = make_struct(ps, arch)
        if not batch.is_cuda:
            raise RuntimeError("t2ms_b200 LA-VAE needs CUDA tensors (no CPU fallback)")
        B, L = batch.shape[0], batch.shape[-1]
        x = batch.detach().to(torch.float32).reshape(B, arch["in_channels"], L).contiguous()
        gst = None
        if backward:
            grads = {}
            for n, p in ps.items():
                if p.requires_grad:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                    grads[n] = p.grad
            scratch = {n: torch.zeros_like(p) for n, p in ps.items() if n not in grads}      # frozen parameters: discarded
            gst = make_struct({**grads, **scratch}, arch)
        recon = torch.empty(B, arch["in_channels"], L, device=x.device, dtype=torch.float32)
        z = torch.empty(B, arch["emb"], self.flow_dim, device=x.device, dtype=torch.float32)
        sums = torch.zeros(2, device=x.device, dtype=torch.float32)
        wp, nb, _ = self._workspace(st, B, L, x.device)
        with torch.cuda.device(x.device):
            rc = _lib.load().t2s_lavae_train_step(C.byref(st), C.byref(gst) if gst is not None else None, x.data_ptr(), recon.data_ptr(),
                                                  z.data_ptr(), sums.data_ptr(), B, L, wp, nb, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "t2s_lavae_train_step")
        recon_error = sums[0] / float(B * arch["in_channels"] * L)
        cross = sums[1] / float(B * arch["emb"] * (L // 4))
        return recon_error + cross, recon_error, recon, z
