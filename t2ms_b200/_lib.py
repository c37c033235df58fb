"""ctypes binding of the C ABI declared in include/t2s_b200.h.

The product path is CUDA only: if the shared library is missing this raises, there is no CPU or
PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# T2S_B200_LIB: another build of the same sources (A/B experiments: tools/build_variant.py); the default is the in-tree build
LIB_PATH = os.environ.get("T2S_B200_LIB") or os.path.join(HERE, "lib", "libt2s_b200.so")

EXPORTS = [
    "t2s_version", "t2s_last_error", "t2s_init", "t2s_debug_set_phase_trace", "t2s_set_fused", "t2s_set_pdl", "t2s_debug_set_fused_stats", "t2s_debug_set_fused_trace", "t2s_dit_workspace_bytes", "t2s_dit_workspace_offsets",
    "t2s_dit_workspace_bytes_h", "t2s_dit_workspace_offsets_h", "t2s_dit_attention_h",
    "t2s_dit_forward", "t2s_sample", "t2s_sample_ddpm_seeded", "t2s_vae_decode", "t2s_vae_encode",
    "t2s_dit_cond", "t2s_dit_embed_qkv", "t2s_dit_attention", "t2s_dit_block_post", "t2s_dit_final",
    "t2s_train_workspace_bytes", "t2s_train_workspace_bytes_h", "t2s_train_make_inputs_h", "t2s_dit_train_step", "t2s_dit_train_forward", "t2s_dit_train_backward",
    "t2s_train_make_inputs", "t2s_adamw_step", "t2s_gemm_tf32",
    "t2s_series_metrics", "t2s_rf_euler", "t2s_ddpm_p_sample", "t2s_lavae_workspace_bytes", "t2s_lavae_encode", "t2s_lavae_decode", "t2s_lavae_train_step", "t2s_train_attention_scratch_bytes", "t2s_train_attention_forward", "t2s_train_attention_backward",
]

P = C.c_void_p


class DitWeights(C.Structure):
    _fields_ = [("w_qkv", P * 4), ("w_post", P * 4), ("b_qkv", P * 4), ("b_proj", P * 4), ("b_fc1", P * 4),
                ("b_fc2", P * 4), ("w_ada_t", P), ("b_ada", P), ("w_embed", P), ("b_embed", P), ("pos", P),
                ("w_final", P), ("b_final", P), ("freqs", P), ("latent_h", C.c_int), ("w_qkv_half", P * 4), ("w_post_half", P * 4)]


class DitParams(C.Structure):
    """t2s_dit_params: raw fp32 parameter (or gradient) pointers in the reference state-dict layouts."""
    _fields_ = [("conv_w", P), ("conv_b", P), ("pe_w", P), ("pe_b", P), ("pos", P), ("ln_w", P), ("ln_b", P),
                ("lf_w", P), ("lf_b", P), ("freqs", P), ("qkv_w", P * 4), ("qkv_b", P * 4), ("proj_w", P * 4),
                ("proj_b", P * 4), ("fc1_w", P * 4), ("fc1_b", P * 4), ("fc2_w", P * 4), ("fc2_b", P * 4),
                ("ada_w", P * 4), ("ada_b", P * 4), ("latent_h", C.c_int)]


class LavaeParams(C.Structure):
    """t2s_lavae_params: architecture ints + raw fp32 parameter (or gradient) pointers in the reference layouts."""
    _fields_ = [("in_channels", C.c_int), ("hidden", C.c_int), ("res_hidden", C.c_int), ("emb", C.c_int), ("n_res", C.c_int),
                ("flow_dim", C.c_int),
                ("enc_conv1_w", P), ("enc_conv1_b", P), ("enc_conv2_w", P), ("enc_conv2_b", P), ("enc_conv3_w", P), ("enc_conv3_b", P),
                ("enc_res_w3", P * 4), ("enc_res_w1", P * 4), ("enc_pre_w", P), ("enc_pre_b", P),
                ("dec_conv1_w", P), ("dec_conv1_b", P), ("dec_res_w3", P * 4), ("dec_res_w1", P * 4),
                ("dec_ct1_w", P), ("dec_ct1_b", P), ("dec_ct2_w", P), ("dec_ct2_b", P)]


class VaeDecWeights(C.Structure):
    _fields_ = [("conv1_w", P), ("conv1_b", P), ("res_w3", P * 2), ("res_w1", P * 2), ("ct1_w", P), ("ct1_b", P),
                ("ct2_w", P), ("ct2_b", P)]


class VaeEncWeights(C.Structure):
    _fields_ = [("conv1_w", P), ("conv1_b", P), ("conv2_w", P), ("conv2_b", P), ("conv3_w", P), ("conv3_b", P),
                ("res_w3", P * 2), ("res_w1", P * 2), ("pre_w", P), ("pre_b", P)]


_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load libt2s_b200.so (built by t2ms_b200.build / __graft_entry__.build)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m t2ms_b200.build` "
                "(t2ms_b200 has no CPU / PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        i, f, sz = C.c_int, C.c_float, C.c_size_t
        lib.t2s_version.restype = i
        lib.t2s_last_error.restype = C.c_char_p
        lib.t2s_init.restype = i
        lib.t2s_debug_set_phase_trace.restype = None
        lib.t2s_debug_set_phase_trace.argtypes = [P]
        lib.t2s_set_pdl.restype = None
        lib.t2s_set_pdl.argtypes = [i]
        lib.t2s_set_fused.restype = None
        lib.t2s_set_fused.argtypes = [i, i]
        lib.t2s_debug_set_fused_stats.restype = None
        lib.t2s_debug_set_fused_stats.argtypes = [P]
        lib.t2s_debug_set_fused_trace.restype = None
        lib.t2s_debug_set_fused_trace.argtypes = [P]
        lib.t2s_dit_workspace_bytes.restype = sz
        lib.t2s_dit_workspace_bytes.argtypes = [i]
        lib.t2s_dit_workspace_offsets.restype = None
        lib.t2s_dit_workspace_offsets.argtypes = [i, C.POINTER(sz * 4)]
        lib.t2s_dit_workspace_bytes_h.restype = sz
        lib.t2s_dit_workspace_bytes_h.argtypes = [i, i]
        lib.t2s_dit_workspace_offsets_h.restype = i
        lib.t2s_dit_workspace_offsets_h.argtypes = [i, i, C.POINTER(sz * 4)]
        lib.t2s_dit_attention_h.restype = i
        lib.t2s_dit_attention_h.argtypes = [i, i, P, P]
        lib.t2s_dit_forward.restype = i
        lib.t2s_dit_forward.argtypes = [C.POINTER(DitWeights), P, P, P, P, i, P, sz, P]
        lib.t2s_sample.restype = i
        lib.t2s_sample.argtypes = [C.POINTER(DitWeights), i, P, P, P, C.POINTER(f), P, P, i, i, f, P, sz, P]
        lib.t2s_sample_ddpm_seeded.restype = i
        lib.t2s_sample_ddpm_seeded.argtypes = [C.POINTER(DitWeights), P, P, P, C.POINTER(f), C.c_ulonglong, P, i, i, f, P, sz, P]
        lib.t2s_vae_decode.restype = i
        lib.t2s_vae_decode.argtypes = [C.POINTER(VaeDecWeights), P, P, P, i, i, P]
        lib.t2s_vae_encode.restype = i
        lib.t2s_vae_encode.argtypes = [C.POINTER(VaeEncWeights), P, P, P, i, i, P]
        lib.t2s_dit_cond.restype = i
        lib.t2s_dit_cond.argtypes = [C.POINTER(DitWeights), P, i, P, i, i, P, P]
        lib.t2s_dit_embed_qkv.restype = i
        lib.t2s_dit_embed_qkv.argtypes = [C.POINTER(DitWeights), P, i, i, P, P]
        lib.t2s_dit_attention.restype = i
        lib.t2s_dit_attention.argtypes = [i, P, P]
        lib.t2s_dit_block_post.restype = i
        lib.t2s_dit_block_post.argtypes = [C.POINTER(DitWeights), i, i, P, P]
        lib.t2s_dit_final.restype = i
        lib.t2s_dit_final.argtypes = [C.POINTER(DitWeights), P, i, P, P]
        d = C.c_double
        lib.t2s_train_workspace_bytes.restype = sz
        lib.t2s_train_workspace_bytes.argtypes = [i]
        lib.t2s_train_workspace_bytes_h.restype = sz
        lib.t2s_train_workspace_bytes_h.argtypes = [i, i]
        lib.t2s_train_make_inputs_h.restype = i
        lib.t2s_train_make_inputs_h.argtypes = [i, P, P, P, P, P, P, i, i, P]
        lib.t2s_dit_train_step.restype = i
        lib.t2s_dit_train_step.argtypes = [C.POINTER(DitParams), C.POINTER(DitParams), P, P, P, P, P, P, i, d, P, sz, P]
        lib.t2s_dit_train_forward.restype = i
        lib.t2s_dit_train_forward.argtypes = [C.POINTER(DitParams), P, P, P, P, i, P, sz, P]
        lib.t2s_dit_train_backward.restype = i
        lib.t2s_dit_train_backward.argtypes = [C.POINTER(DitParams), C.POINTER(DitParams), P, i, P, sz, P]
        lib.t2s_train_make_inputs.restype = i
        lib.t2s_train_make_inputs.argtypes = [i, P, P, P, P, P, P, i, P]
        lib.t2s_adamw_step.restype = i
        lib.t2s_adamw_step.argtypes = [P, P, P, P, sz, i, f, f, f, f, f, f, P]
        lib.t2s_gemm_tf32.restype = i
        lib.t2s_gemm_tf32.argtypes = [P, P, P, P, i, i, i, i, i, i, i, i, i, f, i, P]
        lib.t2s_train_attention_scratch_bytes.restype = sz
        lib.t2s_train_attention_scratch_bytes.argtypes = [i]
        lib.t2s_train_attention_forward.restype = i
        lib.t2s_train_attention_forward.argtypes = [P, P, P, i, P, sz, P]
        lib.t2s_train_attention_backward.restype = i
        lib.t2s_train_attention_backward.argtypes = [P, P, P, P, P, i, P, sz, P]
        lib.t2s_rf_euler.restype = i
        lib.t2s_rf_euler.argtypes = [P, P, f, P, sz, P]
        lib.t2s_ddpm_p_sample.restype = i
        lib.t2s_ddpm_p_sample.argtypes = [P, P, P, P, P, P, P, i, i, P]
        lib.t2s_series_metrics.restype = i
        lib.t2s_series_metrics.argtypes = [P, P, i, i, P, P, P]
        lib.t2s_lavae_workspace_bytes.restype = sz
        lib.t2s_lavae_workspace_bytes.argtypes = [C.POINTER(LavaeParams), i, i]
        lib.t2s_lavae_encode.restype = i
        lib.t2s_lavae_encode.argtypes = [C.POINTER(LavaeParams), P, P, P, i, i, P, sz, P]
        lib.t2s_lavae_decode.restype = i
        lib.t2s_lavae_decode.argtypes = [C.POINTER(LavaeParams), P, P, P, i, i, P, sz, P]
        lib.t2s_lavae_train_step.restype = i
        lib.t2s_lavae_train_step.argtypes = [C.POINTER(LavaeParams), C.POINTER(LavaeParams), P, P, P, P, i, i, P, sz, P]
        _lib = lib
        return lib


def check(rc: int, what: str = "t2s") -> None:
    if rc != 0:
        msg = load().t2s_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
