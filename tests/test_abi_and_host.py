"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/t2s_b200.h declares;
host logic (packing layouts, coefficient tables, sharding, module interfaces) behaves like the reference."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, T, load_golden
from oracle import t2s_oracle as O
from t2ms_b200 import DDPM, RectifiedFlow, Transformer, _lib, synth, vqvae
from t2ms_b200.compat import VAE_ARGS
from t2ms_b200.packing import WSTAGE_BYTES, tile_rows, umma_bias_block, umma_stage, umma_wstage
from t2ms_b200.sampler import gather_series, shard_range


@pytest.fixture(scope="module")
def lib():
    from t2ms_b200.build import build
    build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "t2s_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(t2s_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/t2s_b200.h but not exported"
    assert sorted(declared) == sorted(_lib.EXPORTS)
    assert lib.t2s_version() == 200


def test_workspace_sizes(lib):
    off = (ctypes.c_size_t * 4)()
    lib.t2s_dit_workspace_offsets(2048, ctypes.byref(off))
    off = list(off)
    assert off[0] == 0 and off[1] == 1024 * 8 * 128 * 128 * 4 and all(o % 256 == 0 for o in off)
    assert lib.t2s_dit_workspace_bytes(2048) >= off[3] + 2048 * 4 * 768 * 4
    assert lib.t2s_dit_workspace_bytes(1) < lib.t2s_dit_workspace_bytes(2)


def test_bad_arguments_return_error_codes(lib):
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.t2s_dit_forward(None, None, None, None, None, 0, None, 0, None) == -1
    assert b"bad argument" in lib.t2s_last_error()
    assert lib.t2s_sample(None, 0, None, None, None, None, None, None, 1, 1, 7.0, None, 0, None) == -1
    assert lib.t2s_vae_decode(None, None, None, None, 1, 96, None) == -1


def test_no_cpu_fallback():
    m = Transformer()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(input=torch.zeros(1, 64, 30), t=torch.zeros(1), text_input=None)
    v = vqvae(VAE_ARGS)
    with pytest.raises(RuntimeError, match="CUDA"):
        v.decoder(torch.zeros(1, 64, 30), length=24)
    with pytest.raises(RuntimeError, match="CUDA"):
        v.encoder(torch.zeros(1, 24))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "t2ms_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_umma_stage_layout():
    """Weight stages are tcgen05 no-swizzle K-major operand images: (n,k) at (k//8)*2048 + (n//8)*128 + (n%8)*16 + (k%8)*2 bytes."""
    w = (torch.arange(128 * 128, dtype=torch.float32).reshape(128, 128) % 2048)          # exact in fp16
    img = umma_stage(w)
    assert img.dtype == torch.float16 and img.numel() == 128 * 128
    for n, k in ((0, 0), (1, 0), (7, 9), (8, 8), (77, 127), (127, 64), (127, 127)):
        off = ((k // 8) * 2048 + (n // 8) * 128 + (n % 8) * 16 + (k % 8) * 2) // 2
        assert float(img[off]) == float(w[n, k])
    pos = torch.arange(480 * 128, dtype=torch.float32).reshape(480, 128)
    tiled = tile_rows(pos)
    assert tiled.shape == (8, 32, 64, 4)
    assert torch.equal(tiled[3, 5, 17], pos[3 * 60 + 17, 20:24]) and float(tiled[:, :, 60:].abs().sum()) == 0.0


def test_weight_stage_carries_the_bias_block():
    """A token_kernel weight stage = the 32 KB operand image + a [128 n][16 k] bias block in the same layout: k = 0 / 1 hold the
    fp16 high / low parts of the fp32 bias (their sum is the bias to 2^-22), everything else is zero; fc2's second K half and
    out-of-range biases are handled (include/t2s_b200.h: t2s_dit_weights.w_qkv / w_post)."""
    g = torch.Generator().manual_seed(3)
    w, b = torch.randn(128, 128, generator=g), torch.randn(128, generator=g) * 2.0
    st = umma_wstage(w, b)
    assert st.dtype == torch.float16 and st.numel() * 2 == WSTAGE_BYTES
    assert torch.equal(st[:128 * 128], umma_stage(w))
    blk = st[128 * 128:].float()
    for n in (0, 1, 7, 8, 77, 127):
        base = ((n // 8) * 128 + (n % 8) * 16) // 2                       # K chunk 0
        assert abs(float(blk[base] + blk[base + 1]) - float(b[n])) <= abs(float(b[n])) * 2.0 ** -21 + 1e-9
        assert float(blk[base]) == float(b[n].to(torch.float16)) and float(blk[base + 2:base + 8].abs().sum()) == 0.0
    assert float(blk[1024:].abs().sum()) == 0.0                           # K chunk 1 (k = 8..15) is zero
    assert float(umma_bias_block(None).abs().sum()) == 0.0 and umma_bias_block(None).numel() == 2048
    with pytest.raises(RuntimeError):
        umma_bias_block(torch.full((128,), 1.0e5))


def test_module_interfaces_match_reference_state_dict():
    sd = synth.make_dit_state(3)
    m = Transformer()
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape, k
    m.load_state_dict(sd, strict=True)
    assert not m.pos_embed.requires_grad
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 946265
    # default init: adaLN zero, biases zero (transformer.py:194-204)
    m2 = Transformer()
    assert float(m2.layers[0].adaLN_modulation[-1].weight.abs().sum()) == 0.0
    assert float(m2.layers[2].attn.qkv.bias.abs().sum()) == 0.0
    v = vqvae(VAE_ARGS)
    vsd = synth.make_vae_state(4)
    assert sorted(v.state_dict().keys()) == sorted(vsd.keys())
    assert sum(p.numel() for p in v.parameters()) == 672833
    # attaching the frozen encoder like train.py:30-33 / infer.py:47
    m.encoder = v.encoder
    names = [n for n, _ in m.named_parameters()]
    assert sum("encoder" in n for n in names) == 12
    assert len(m.state_dict()) == 67
    with pytest.raises(ValueError):
        vqvae(type(VAE_ARGS)(block_hidden_size=64, num_residual_layers=2, res_hidden_size=256, embedding_dim=64))


def test_backbone_classes_host_side():
    """The schedule tables equal the reference's (golden from the reference classes); the per-step methods run on the GPU
    through the custom ops (tests/test_gpu_outputs.py::test_backbone_classes_on_cuda_match_reference_golden) and refuse CPU
    tensors like every other entry point; the losses stay torch ops (differentiable)."""
    g = load_golden("backbone.npz")
    x1, eps, ti = (T(g[k]) for k in ("x1", "eps", "ti"))
    rf, dd = RectifiedFlow(), DDPM(1000, "cpu")
    for a, k in ((dd.beta, "beta"), (dd.alpha, "alpha"), (dd.alpha_bar, "alpha_bar")):
        assert np.array_equal(a.numpy(), g[k])
    assert dd.sigma2 is dd.beta and dd.total_steps == 1000
    assert abs(float(rf.loss(x1, eps)) - float(g["rf_loss"])) < 1e-6 and abs(float(dd.loss(x1, eps)) - float(g["ddpm_loss"])) < 1e-6
    for call in (lambda: rf.euler(x1, eps, 0.01), lambda: rf.create_flow(x1, T(g["t"])), lambda: dd.q_sample(x1, ti, eps),
                 lambda: dd.p_sample(x1, eps, ti)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_custom_ops_are_registered_with_fake_implementations():
    """north_star: "a thin C-ABI PyTorch custom-op layer".  Every entry point the modules use is a torch.library op with a
    fake (meta) implementation: shapes propagate under FakeTensorMode without touching a GPU or the shared library."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from t2ms_b200 import ops
    for name in ops.OPS:
        assert hasattr(torch.ops.t2s_b200, name), name
    with FakeTensorMode():
        x = torch.empty(5, 64, 30, device="cuda")
        t, emb, ws = torch.empty(5, device="cuda"), torch.empty(5, 128, device="cuda"), torch.empty(1024, dtype=torch.uint8, device="cuda")
        assert torch.ops.t2s_b200.dit_forward(x, t, emb, ws, 1, 30).shape == (5, 64, 30)
        assert torch.ops.t2s_b200.dit_forward(x, t, None, ws, 1, 30).shape == (5, 64, 30)
        s, after = torch.ops.t2s_b200.vae_decode(x, 96, 1)
        assert s.shape == (5, 96) and after.shape == (5, 64, 24)
        z, before = torch.ops.t2s_b200.vae_encode(torch.empty(5, 48, device="cuda"), 1)
        assert z.shape == (5, 64, 30) and before.shape == (5, 64, 12)
        assert torch.ops.t2s_b200.rf_euler(x, x, 0.01).shape == x.shape
        c = torch.empty(5, device="cuda")
        assert torch.ops.t2s_b200.ddpm_p_sample(x, x, x, c, c, c).shape == x.shape
        xt, tg = torch.ops.t2s_b200.make_inputs(0, x, x, c, None, 30)
        assert xt.shape == x.shape and tg.shape == x.shape
        assert torch.ops.t2s_b200.sample_loop(x, emb, t, torch.empty(4, 3), None, None, ws, 0, 4, 7.0, 0, False, 1, 30) is None


def test_sampler_tables_match_oracle_formulas():
    for steps in (4, 10, 100):
        tt = RectifiedFlow.timesteps(steps)
        assert torch.equal(tt, O.rf_timesteps(steps, 1)[:, 0])
        c = RectifiedFlow.coefficients(steps)
        assert c[0, 0].item() == np.float32(1.0 / steps) and float(c[:, 1:].abs().sum()) == 0
    steps = 50
    t = DDPM.timesteps(steps)
    assert t.tolist() == [math.floor(steps - 1 - j) for j in range(steps)]
    coef = DDPM.coefficients(steps)
    sched = O.ddpm_schedule(steps)
    x, e, n = torch.randn(3, 64, 30), torch.randn(3, 64, 30), torch.randn(3, 64, 30)
    for j in (0, 7, 49):
        tj = torch.full((3,), int(t[j]), dtype=torch.long)
        ref = O.ddpm_p_sample(x, e, tj, n, sched)
        mine = coef[j, 0] * (x - coef[j, 1] * e) + coef[j, 2] * n
        assert torch.allclose(mine, ref, atol=1e-6, rtol=1e-6)


def test_shard_range_partitions():
    for total in (1, 7, 512, 1024, 1025):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from t2ms_b200.sampler import gather_series, shard_range
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
total = 7
full = torch.arange(total * 5, dtype=torch.float32).reshape(total, 5)
lo, hi = shard_range(total, rank, world)
out = gather_series(full[lo:hi].clone(), total)
assert torch.equal(out, full), (rank, out)
dist.destroy_process_group()
print("ok", rank)
'''


def test_gather_series_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_compat_aliases():
    from t2ms_b200 import compat
    compat.install()
    from model.denoiser.transformer import Transformer as RT
    from model.backbone.DDPM import DDPM as RD
    from model.pretrained.vqvae import vqvae as RV
    assert RT is Transformer and RD is DDPM and RV is vqvae
    v = compat.convert_reference_vae(synth.make_vae_state(5))
    assert isinstance(v, vqvae)


_COMPAT_CHECKOUT = r'''
import sys
sys.path.insert(0, sys.argv[1])          # the repo
sys.path.insert(0, sys.argv[2])          # a reference-style checkout: model/ is a real (namespace) package with other submodules
from t2ms_b200 import compat
compat.install()                         # INTEGRATION.md 2(a): first line of an untouched infer.py / train.py
# the import lines of infer.py:4-8 / train.py:6-11
from model.denoiser.mlp import MLP
from model.denoiser.transformer import Transformer
from model.backbone.rectified_flow import RectifiedFlow
from model.backbone.DDPM import DDPM
import t2ms_b200
assert MLP.marker == "real reference submodule"
assert Transformer is t2ms_b200.Transformer and RectifiedFlow is t2ms_b200.RectifiedFlow and DDPM is t2ms_b200.DDPM
import model.pretrained.vqvae as V
assert V.vqvae is t2ms_b200.vqvae          # the class path the pickled LA-VAE resolves (infer.py:39-41)
print("ok")
'''


def test_compat_install_keeps_real_reference_submodules_importable(tmp_path):
    """install() must not hide the reference's other submodules (model.denoiser.mlp, infer.py:4) behind empty
    fake packages when the scripts run from a reference checkout."""
    co = tmp_path / "checkout"
    (co / "model" / "denoiser").mkdir(parents=True)
    (co / "model" / "backbone").mkdir()
    (co / "model" / "pretrained").mkdir()
    (co / "model" / "denoiser" / "mlp.py").write_text("class MLP:\n    marker = 'real reference submodule'\n")
    (co / "model" / "denoiser" / "transformer.py").write_text("raise ImportError('the reference denoiser must not be imported')\n")
    script = tmp_path / "c.py"
    script.write_text(_COMPAT_CHECKOUT)
    r = subprocess.run([sys.executable, str(script), ROOT, str(co)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_fork_module_aliases_and_state_dicts():
    """The fork's module paths (mytrain.py:7-9, myinfer.py:5-6) resolve to the t2ms_b200 classes; Transformer(dim) and the
    multivariate vqvae(args) carry the reference's parameter names / shapes (strict load of synthetic reference-shaped
    state dicts) on CPU — only their forward needs the GPU."""
    from argparse import Namespace
    from t2ms_b200 import compat
    compat.install()
    from model.denoiser.mytransformer import Transformer as FT
    from model.pretrained.myvqvae import vqvae as FV
    from t2ms_b200 import mylavae
    assert FT is Transformer and FV is mylavae.vqvae
    for dim in (50, 64):
        m = FT(dim)
        m.load_state_dict(synth.make_dit_state(3, dim=dim), strict=True)
        assert m.pos_embed.shape == (1, 16 * dim, 128) and m.patch_count == 16 * dim
    with pytest.raises(ValueError):
        FT(48)
    v = FV(Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256, embedding_dim=64, flow_dim=50, input_dim=7))
    v.load_state_dict(synth.make_vae_state(4, in_channels=7), strict=True)
    assert v.encoder._conv_1.weight.shape == (64, 7, 4) and v.decoder._conv_trans_2.weight.shape == (64, 7, 4)
    with pytest.raises(RuntimeError):
        v.encoder(torch.rand(2, 7, 100))                       # CPU tensors: no fallback


def test_tools_compile():
    """The measurement tools the profiles were produced with stay syntactically valid, and the probes that include the product
    kernels (tools/probe_pass.cu) still compile against them."""
    import py_compile
    tools = os.path.join(ROOT, "tools")
    for f in sorted(os.listdir(tools)):
        if f.endswith(".py"):
            py_compile.compile(os.path.join(tools, f), doraise=True)
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-I", os.path.join(ROOT, "t2ms_b200", "csrc"),
                        "-c", "-o", os.devnull, os.path.join(tools, "probe_pass.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
