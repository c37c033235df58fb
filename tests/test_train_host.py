"""CPU-side checks of the training path: parameter bookkeeping, C-ABI argument validation (no compute without a
GPU), the no-fallback rule, and the data-parallel gradient convention (world-size-2 gloo, oracle arithmetic)."""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden
from t2ms_b200 import Transformer, _lib, synth
from t2ms_b200.training import FlatBuffer, trainable_names, _struct


@pytest.fixture(scope="module")
def lib():
    from t2ms_b200.build import build
    build()
    return _lib.load()


def test_trainable_names_match_reference_gradients():
    g = load_golden("train.npz")
    ref = sorted(str(n).replace("grad_norm/", "") for n in g["names"])      # parameters with a .grad after loss.backward()
    assert sorted(trainable_names()) == ref and len(ref) == 48
    m = Transformer()
    named = dict(m.named_parameters())
    assert sum(named[n].numel() for n in trainable_names()) == 925_592       # the all-reduce payload


def test_flat_buffer_layout_and_param_struct():
    m = Transformer()
    shapes = {n: p.shape for n, p in m.named_parameters() if n in set(trainable_names())}
    fb = FlatBuffer(shapes, "cpu")
    for n, (off, s) in fb.offsets.items():
        assert off % 64 == 0 and fb.view(n).shape == s
    assert fb.flat.numel() >= 925_592
    st = _struct(fb.views(), None, None)
    assert ctypes.sizeof(_lib.DitParams) == (10 + 10 * 4) * 8 + 8      # 50 pointers + latent_h (padded)
    assert st.qkv_w[3] == fb.view("layers.3.attn.qkv.weight").data_ptr() and st.lf_b == fb.view("linear_emb_to_patch.bias").data_ptr()


def test_training_entry_points_validate_arguments(lib):
    assert lib.t2s_dit_train_step(None, None, None, None, None, None, None, None, 0, 1.0, None, 0, None) == -1
    assert b"bad argument" in lib.t2s_last_error()
    assert lib.t2s_dit_train_forward(None, None, None, None, None, 1, None, 0, None) == -1
    assert lib.t2s_dit_train_backward(None, None, None, 1, None, 0, None) == -1
    assert lib.t2s_gemm_tf32(None, None, None, None, 1, 1, 1, 4, 4, 4, 0, 0, 0, 1.0, 1, None) == -1
    assert lib.t2s_adamw_step(None, None, None, None, 0, 1, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1.0, None) == -1
    assert lib.t2s_train_make_inputs(2, None, None, None, None, None, None, 1, None) == -1
    assert lib.t2s_train_workspace_bytes(2) > lib.t2s_train_workspace_bytes(1) > 480 * 7000 * 4


def test_training_has_no_cpu_fallback():
    m = Transformer().train()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(input=torch.zeros(2, 64, 30), t=torch.zeros(2), text_input=torch.zeros(2, 128))
    from t2ms_b200.training import DitTrainer
    with pytest.raises(RuntimeError, match="CUDA"):
        DitTrainer(m)


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import t2s_oracle as O
from t2ms_b200 import synth
from t2ms_b200.sampler import shard_range
torch.set_num_threads(2)
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
B = 4
dsd = synth.make_dit_state(15, bias_std=0.02)
x1, x0, emb = synth.make_noise(B, seed=5), synth.make_noise(B, seed=6), synth.make_text_embeddings(B, seed=7)
t = torch.tensor([0.1, 0.5, 0.73, 1.0])
x_t, target = O.rf_create_flow(x1, t, x0), x1 - x0
_, full = O.train_step_grads(dsd, x_t, t, emb, target)
lo, hi = shard_range(B, rank, world)
# DitTrainer.step convention: every rank normalises by the GLOBAL element count, then one SUM all-reduce
loss_local, g = O.train_step_grads(dsd, x_t[lo:hi], t[lo:hi], emb[lo:hi], target[lo:hi])
scale = (hi - lo) / B
for n in g:
    buf = g[n] * scale
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    err = ((buf - full[n]).norm() / full[n].norm()).item()
    assert err < 1e-5, (n, err)
# the shared classifier-free-guidance coin (train.py:80-82): every rank draws it from an identically seeded CPU generator
# (DitTrainer(coin_seed=...)), whatever its global RNG state is: no broadcast, no device round trip
torch.manual_seed(100 + rank)
gen = torch.Generator().manual_seed(0)
coins = torch.cat([torch.rand(1, generator=gen) for _ in range(16)])
got = [torch.empty_like(coins) for _ in range(world)]
dist.all_gather(got, coins)
assert all(torch.equal(g, got[0]) for g in got)
# ONE all_reduce carries the gradients and the loss sum (the bucket's tail)
flat = torch.cat([torch.full((8,), float(rank + 1)), torch.tensor([0.5 * (rank + 1)])])
dist.all_reduce(flat, op=dist.ReduceOp.SUM)
assert flat[0].item() == 3.0 and flat[-1].item() == 1.5
dist.destroy_process_group()
print("ok", rank)
'''


def test_data_parallel_gradient_convention_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_one_cycle_lr_matches_torch_scheduler():
    from t2ms_b200.train_loop import one_cycle_lr
    for total in (10, 57, 1000):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=1e-4, weight_decay=0.0)
        sch = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-4, total_steps=total)      # train.py:38
        for k in range(total):
            assert abs(one_cycle_lr(k, total) - sch.get_last_lr()[0]) < 1e-12, (total, k)
            opt.step()
            if k < total - 1:
                sch.step()


def test_collate_by_length_groups_like_the_reference():
    import numpy as np
    from t2ms_b200.train_loop import collate_by_length
    rng = np.random.default_rng(0)
    items = []
    for i, (idx, L) in enumerate([(2, 96), (0, 24), (1, 48), (0, 24), (2, 96), (2, 96)]):
        items.append(((np.array([i]), rng.random(L, dtype=np.float32), rng.random(128, dtype=np.float32)), idx))
    out = collate_by_length(items)
    assert [tuple(x.shape) for _, x, _ in out] == [(2, 24), (1, 48), (3, 96)]
    assert [int(t[0]) for t in out[2][0]] == [0, 4, 5] and out[2][2].shape == (3, 128)
    only = collate_by_length(items[:1])
    assert len(only) == 1 and only[0][1].shape == (1, 96)
