"""Helpers shared by the -m gpu parity tests: drive single kernels through the C ABI and view the workspace."""
import ctypes as C

import torch

from t2ms_b200 import _lib, synth
from t2ms_b200.compat import VAE_ARGS
from t2ms_b200.denoiser import Transformer, _aligned
from t2ms_b200.lavae import vqvae

DEV = "cuda:0"


def make_dit(seed, bias_std=0.02):
    sd = synth.make_dit_state(seed, bias_std=bias_std)
    m = Transformer()
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


def make_vae(seed):
    sd = synth.make_vae_state(seed)
    m = vqvae(VAE_ARGS)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


class Workspace:
    """Typed views of the DiT scratch buffer (layout: t2s_dit_workspace_offsets)."""

    def __init__(self, model, nseq):
        lib = _lib.load()
        self.nseq = nseq
        self.buf = model.workspace(nseq, torch.device(DEV))
        self.ptr = _aligned(self.buf)
        self.nbytes = lib.t2s_dit_workspace_bytes(nseq)
        off = (C.c_size_t * 4)()
        lib.t2s_dit_workspace_offsets(nseq, C.byref(off))
        self.off = list(off)
        self.base = self.ptr - self.buf.data_ptr()

    def _view(self, i, nbytes, dtype, shape):
        s = self.base + self.off[i]
        return self.buf[s:s + nbytes].view(dtype).view(shape)

    def _tiles(self, raw):
        """(npair, 8 tiles, 128 rows, 128) tile rows -> (nseq, 480, 128): row = branch*64 + token%60."""
        npair = (self.nseq + 1) // 2
        t = raw.reshape(npair, 8, 2, 64, 128)[:, :, :, :60]                  # pair, tile, branch, tl, feat
        t = t.permute(0, 2, 1, 3, 4).reshape(npair * 2, 480, 128)
        return t[: self.nseq]

    def h(self):
        npair = (self.nseq + 1) // 2
        raw = self._view(0, npair * 8 * 128 * 128 * 4, torch.float32, (npair, 8, 32, 128, 4))   # [c4][row][4]
        return self._tiles(raw.permute(0, 1, 3, 2, 4).reshape(npair, 8, 128, 128))

    def qkv(self):
        """-> (nseq, 3, 4 heads, 480, 32) fp32 from the per-(sequence, head) tcgen05 operand images."""
        n = self.nseq
        raw = self._view(1, n * 4 * 47104 * 2, torch.float16, (n, 4, 47104))
        q = raw[:, :, :16384].reshape(n, 4, 4, 4, 128, 8)[:, :, :, :, :120]          # [qt][d/8][row][8]
        q = q.permute(0, 1, 2, 4, 3, 5).reshape(n, 4, 480, 32)
        k = raw[:, :, 16384:31744].reshape(n, 4, 4, 480, 8).permute(0, 1, 3, 2, 4).reshape(n, 4, 480, 32)   # [d/8][key][8]
        v = raw[:, :, 31744:].reshape(n, 4, 60, 4, 8, 8).permute(0, 1, 2, 4, 3, 5).reshape(n, 4, 480, 32)   # [key/8][d/8][key%8][8]
        return torch.stack([q, k, v], dim=1).float()

    def o(self):
        npair = (self.nseq + 1) // 2
        raw = self._view(2, npair * 8 * 128 * 128 * 2, torch.float16, (npair, 8, 16, 16, 8, 8))  # [k//8][r//8][r%8][k%8]
        return self._tiles(raw.permute(0, 1, 3, 4, 2, 5).reshape(npair, 8, 128, 128)).float()

    def mod(self):
        """adaLN table as the reference lays it out (the kernels keep 1 + scale in the two scale slices)."""
        m = self._view(3, self.nseq * 4 * 768 * 4, torch.float32, (self.nseq, 4, 768)).clone()
        m[..., 128:256] -= 1.0
        m[..., 512:640] -= 1.0
        return m


def stream():
    return torch.cuda.current_stream().cuda_stream


def pack_qkv_images(q, k, v):
    """(nseq, 4 heads, 480, 32) fp32 x3 -> the per-(sequence, head) fp16 operand images attn_kernel reads
    (inverse of Workspace.qkv): Q [q-tile][d/8][128 rows][8], K [d/8][key][8], V [key/8][d/8][key%8][8]."""
    n = q.shape[0]
    qi = torch.zeros(n, 4, 4, 4, 128, 8, dtype=torch.float16)
    qi[:, :, :, :, :120] = q.to(torch.float16).reshape(n, 4, 4, 120, 4, 8).permute(0, 1, 2, 4, 3, 5)
    ki = k.to(torch.float16).reshape(n, 4, 480, 4, 8).permute(0, 1, 3, 2, 4)
    vi = v.to(torch.float16).reshape(n, 4, 60, 8, 4, 8).permute(0, 1, 2, 4, 3, 5)
    return torch.cat([qi.reshape(n, 4, -1), ki.reshape(n, 4, -1), vi.reshape(n, 4, -1)], dim=2).contiguous()
