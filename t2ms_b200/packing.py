"""Host-side weight packing: reference state-dict tensors -> the device layouts the kernels read.

Layouts are documented in include/t2s_b200.h and DESIGN.md.  Everything here runs once per weight
update (module construction, load_state_dict, optimizer step), never inside the sampling loop.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib

D = 128
WSTAGE_BYTES = 32768 + 4096      # token_kernel weight stage: operand image + bias block (csrc/dit_kernels.cuh)


def umma_stage(w: torch.Tensor) -> torch.Tensor:
    """[128 n][128 k] weight block (nn.Linear layout: row n = output feature, K contiguous) -> fp16
    operand image in the tcgen05 no-swizzle K-major canonical layout: 8x8 core matrices (8 rows x 16 B,
    contiguous 128 B), element (n,k) at byte (k//8)*2048 + (n//8)*128 + (n%8)*16 + (k%8)*2
    (t2ms_b200/csrc/dit_kernels.cuh: tc_gemm, umma_desc with LBO=2048, SBO=128)."""
    assert tuple(w.shape) == (128, 128)
    w16 = w.detach().to(torch.float16).reshape(16, 8, 16, 8)            # [n//8][n%8][k//8][k%8]
    return w16.permute(2, 0, 1, 3).contiguous().reshape(-1)              # [k//8][n//8][n%8][k%8]


def umma_bias_block(b: torch.Tensor | None) -> torch.Tensor:
    """fp32 bias [128] -> the 4 KB fp16 operand block [128 n][16 k] that follows a weight image in a stage (same canonical
    layout, two K chunks): k = 0 holds fp16(b), k = 1 holds fp16(b - fp16(b)), everything else is zero.  token_kernel multiplies
    it with a constant A block whose k = 0, 1 columns are 1.0, so the GEMM's fp32 accumulator starts from the bias to 2^-22
    relative (t2ms_b200/csrc/dit_kernels.cuh: tc_gemm).  None -> a zero block (fc2's second K half)."""
    blk = torch.zeros(2, 16, 8, 8, dtype=torch.float16, device=b.device if b is not None else None)   # [k//8][n//8][n%8][k%8]
    if b is not None:
        b = b.detach().to(torch.float32).reshape(16, 8)
        if not bool((b.abs() < 6.0e4).all()):
            raise RuntimeError("t2ms_b200: a Linear bias is outside the fp16 operand range (|b| < 6e4) of the sampling kernels")
        hi = b.to(torch.float16)
        lo = (b - hi.to(torch.float32)).to(torch.float16)
        blk[0, :, :, 0], blk[0, :, :, 1] = hi, lo
    return blk.reshape(-1)


def umma_wstage(w: torch.Tensor, b: torch.Tensor | None) -> torch.Tensor:
    """One token_kernel weight stage (36 KB): the weight image followed by the bias block."""
    return torch.cat([umma_stage(w), umma_bias_block(b).to(w.device)])


def umma_half_stages(w: torch.Tensor) -> torch.Tensor:
    """[128 n][128 k] weight block -> two fp16 HALF-stage images [64 n][128 k] (rows 0..63, rows 64..127), each in the
    tcgen05 no-swizzle K-major canonical layout with a K-chunk stride of 1024 B: element (n,k) of a half at byte
    (k//8)*1024 + (n//8)*128 + (n%8)*16 + (k%8)*2 (t2ms_b200/csrc/dit_fused.cuh: fs_gemm_half, LBO=1024, SBO=128)."""
    assert tuple(w.shape) == (128, 128)
    w16 = w.detach().to(torch.float16).reshape(2, 8, 8, 16, 8)          # [half][n//8][n%8][k//8][k%8]
    return w16.permute(0, 3, 1, 2, 4).contiguous().reshape(-1)            # [half][k//8][n//8][n%8][k%8]


TILE_TOK = {30: 60, 50: 50, 64: 64}      # tokens per pair tile for latent width H (csrc/common.cuh: DitShape)


def tile_rows(t: torch.Tensor, tile_tok: int = 60) -> torch.Tensor:
    """[tokens][128] fp32 -> [tiles][32 col chunks][64 rows][4] (tile_tok valid rows per tile): the
    residual-stream tile layout, in which a warp's 32 rows x 16 B of one column chunk are contiguous."""
    ntile = t.shape[0] // tile_tok
    assert ntile * tile_tok == t.shape[0]
    out = torch.zeros(ntile, 32, 64, 4, dtype=t.dtype, device=t.device)
    out[:, :, :tile_tok, :] = t.reshape(ntile, tile_tok, 32, 4).permute(0, 2, 1, 3)
    return out.contiguous()


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


class PackedDit:
    """Device-resident packed DiT weights + the ctypes struct handed to the C ABI."""

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device):
        f32 = dict(device=device, dtype=torch.float32)
        g = lambda k: sd[k].detach().to(**f32)
        keep = []
        st = _lib.DitWeights()
        for l in range(4):
            p = f"layers.{l}."
            wq = g(p + "attn.qkv.weight")                                           # [384][128]
            bq, bp = g(p + "attn.qkv.bias").contiguous(), g(p + "attn.proj.bias").contiguous()
            b1, b2 = g(p + "mlp.fc1.bias").contiguous(), g(p + "mlp.fc2.bias").contiguous()
            qkv = torch.cat([umma_wstage(wq[i * D:(i + 1) * D], bq[i * D:(i + 1) * D]) for i in range(3)])
            w1, w2 = g(p + "mlp.fc1.weight"), g(p + "mlp.fc2.weight")               # [256][128], [128][256]
            post = [umma_wstage(g(p + "attn.proj.weight"), bp), umma_wstage(w1[:D], b1[:D]), umma_wstage(w1[D:], b1[D:]),
                    umma_wstage(w2[:, :D].contiguous(), b2), umma_wstage(w2[:, D:].contiguous(), None)]
            post = torch.cat(post)
            assert qkv.numel() * 2 == 3 * WSTAGE_BYTES and post.numel() * 2 == 5 * WSTAGE_BYTES
            if sd["pos_embed"].shape[-2] == 480:                                    # T2S shape: half stages for the fused step kernel
                qkv_h = torch.cat([umma_half_stages(wq[i * D:(i + 1) * D]) for i in range(3)])
                f2a, f2b = umma_half_stages(w2[:, :D].contiguous()).reshape(2, -1), umma_half_stages(w2[:, D:].contiguous()).reshape(2, -1)
                post_h = torch.cat([umma_half_stages(g(p + "attn.proj.weight")), umma_half_stages(w1[:D]), umma_half_stages(w1[D:]),
                                    f2a[0], f2b[0], f2a[1], f2b[1]])
                assert qkv_h.numel() * 2 == 6 * 16384 and post_h.numel() * 2 == 10 * 16384
                keep += [qkv_h, post_h]
                st.w_qkv_half[l], st.w_post_half[l] = _ptr(qkv_h), _ptr(post_h)
            keep += [qkv, post, bq, bp, b1, b2]
            st.w_qkv[l], st.w_post[l] = _ptr(qkv), _ptr(post)
            st.b_qkv[l], st.b_proj[l], st.b_fc1[l], st.b_fc2[l] = _ptr(bq), _ptr(bp), _ptr(b1), _ptr(b2)
        w_ada_t = torch.stack([g(f"layers.{l}.adaLN_modulation.1.weight").t().contiguous() for l in range(4)]).contiguous()
        b_ada = torch.stack([g(f"layers.{l}.adaLN_modulation.1.bias") for l in range(4)]).contiguous()
        wpe, wc = g("patch_emb.weight"), g("conv.weight").reshape(4, 4)              # conv [oc][p*2+q]
        w_embed = (wpe @ wc).t().contiguous()                                        # [4][128] (pixel-major)
        b_embed = (wpe @ g("conv.bias") + g("patch_emb.bias")).contiguous()
        ntok = sd["pos_embed"].shape[-2]
        latent_h = ntok // 16
        if latent_h not in TILE_TOK or ntok != 16 * latent_h:
            raise RuntimeError(f"t2ms_b200 kernels are built for latent widths {sorted(TILE_TOK)} (pos_embed has {ntok} tokens)")
        pos = tile_rows(g("pos_embed").reshape(ntok, D), TILE_TOK[latent_h])
        st.latent_h = latent_h
        wl = g("linear_emb_to_patch.weight")
        w_final = (wl * g("ln.weight").unsqueeze(0)).contiguous()                    # [4][128]
        b_final = (wl @ g("ln.bias") + g("linear_emb_to_patch.bias")).contiguous()
        freqs = torch.pow(10000, torch.linspace(0, 1, D // 2)).to(**f32).contiguous()  # transformer.py:34
        keep += [w_ada_t, b_ada, w_embed, b_embed, pos, w_final, b_final, freqs]
        st.w_ada_t, st.b_ada, st.w_embed, st.b_embed = _ptr(w_ada_t), _ptr(b_ada), _ptr(w_embed), _ptr(b_embed)
        st.pos, st.w_final, st.b_final, st.freqs = _ptr(pos), _ptr(w_final), _ptr(b_final), _ptr(freqs)
        self.struct = st
        self.ref = C.byref(st)
        self._keep = keep
        self.device = device
        self.latent_h = latent_h
        from . import ops
        self.handle = ops.register_handle(self)          # what the torch.library ops take instead of a pointer struct


def _conv_w(w):      # Conv1d weight [oc][ic][k] -> [ic][k][oc]
    return w.permute(1, 2, 0).contiguous()


def _convT_w(w):     # ConvTranspose1d weight [ic][oc][k] -> [ic][k][oc]
    return w.permute(0, 2, 1).contiguous()


class PackedVaeDecoder:
    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device, prefix: str = ""):
        g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)
        st = _lib.VaeDecWeights()
        t = {
            "conv1_w": _conv_w(g("_conv_1.weight")), "conv1_b": g("_conv_1.bias").contiguous(),
            "ct1_w": _convT_w(g("_conv_trans_1.weight")), "ct1_b": g("_conv_trans_1.bias").contiguous(),
            "ct2_w": g("_conv_trans_2.weight").reshape(64, 4).contiguous(), "ct2_b": g("_conv_trans_2.bias").contiguous(),
        }
        assert t["conv1_w"].shape == (64, 3, 128) and t["ct1_w"].shape == (128, 4, 64), \
            "LA-VAE kernels are built for block_hidden_size=128, res_hidden_size=256, embedding_dim=64"
        for k, v in t.items():
            setattr(st, k, _ptr(v))
        for i in range(2):
            w3 = _conv_w(g(f"_residual_stack._layers.{i}._block.1.weight"))
            w1 = g(f"_residual_stack._layers.{i}._block.3.weight").reshape(128, 256).t().contiguous()
            assert w3.shape == (128, 3, 256)
            t[f"w3{i}"], t[f"w1{i}"] = w3, w1
            st.res_w3[i], st.res_w1[i] = _ptr(w3), _ptr(w1)
        self.struct, self.ref, self._keep, self.device = st, C.byref(st), t, device
        from . import ops
        self.handle = ops.register_handle(self)


class PackedVaeEncoder:
    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device, prefix: str = ""):
        g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)
        st = _lib.VaeEncWeights()
        t = {
            "conv1_w": _conv_w(g("_conv_1.weight")), "conv1_b": g("_conv_1.bias").contiguous(),
            "conv2_w": _conv_w(g("_conv_2.weight")), "conv2_b": g("_conv_2.bias").contiguous(),
            "conv3_w": _conv_w(g("_conv_3.weight")), "conv3_b": g("_conv_3.bias").contiguous(),
            "pre_w": g("_pre_vq_conv.weight").reshape(64, 128).t().contiguous(), "pre_b": g("_pre_vq_conv.bias").contiguous(),
        }
        assert t["conv1_w"].shape == (1, 4, 64) and t["conv2_w"].shape == (64, 4, 128) and t["conv3_w"].shape == (128, 3, 128), \
            "LA-VAE kernels are built for block_hidden_size=128, res_hidden_size=256, embedding_dim=64"
        for k, v in t.items():
            setattr(st, k, _ptr(v))
        for i in range(2):
            w3 = _conv_w(g(f"_residual_stack._layers.{i}._block.1.weight"))
            w1 = g(f"_residual_stack._layers.{i}._block.3.weight").reshape(128, 256).t().contiguous()
            t[f"w3{i}"], t[f"w1{i}"] = w3, w1
            st.res_w3[i], st.res_w1[i] = _ptr(w3), _ptr(w1)
        self.struct, self.ref, self._keep, self.device = st, C.byref(st), t, device
        from . import ops
        self.handle = ops.register_handle(self)
