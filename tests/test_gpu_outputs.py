"""GPU parity of the step after the hot path: on-device MSE / WAPE (evaluation.py:166-206) and the generation file
formats of infer.py:100-123, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import make_dit, make_vae
from oracle import t2s_oracle as O
from t2ms_b200 import T2SSampler, load_generation, run_inference, series_metrics, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_series_metrics_match_reference_golden(case):
    """fp32 per-sample sums, fp64 means: relative 2e-6 against the reference's (float32) numpy loops."""
    g = load_golden("eval.npz")
    ori, gen = torch.from_numpy(g[f"{case}/ori"]).to(DEV), torch.from_numpy(g[f"{case}/gen"]).to(DEV)
    m = series_metrics(ori, gen)
    assert abs(m["MSE"] - float(g[f"{case}/mse"])) <= 2e-6 * float(g[f"{case}/mse"])
    assert abs(m["WAPE"] - float(g[f"{case}/wape"])) <= 2e-6 * float(g[f"{case}/wape"])
    assert m["valid"] == ori.shape[0] - (1 if case == "c" else 0)
    # the (N, 1, L) layout of evaluation.py:295-296 gives the same numbers
    m2 = series_metrics(ori.transpose(1, 2), gen.transpose(1, 2))
    assert m2 == m


def test_series_metrics_edge_cases():
    z = torch.zeros(4, 24, device=DEV)
    m = series_metrics(z, z + 1.0)
    assert m["MSE"] == 1.0 and np.isnan(m["WAPE"]) and m["valid"] == 0          # np.nanmean of all-NaN
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(100_000, 96, generator=g), torch.rand(100_000, 96, generator=g)
    m = series_metrics(a.to(DEV), b.to(DEV))
    ref_mse, ref_wape = O.calculate_mse(a.numpy()[:, None, :], b.numpy()[:, None, :]), O.calculate_wape(a.numpy()[:, None, :], b.numpy()[:, None, :])
    assert abs(m["MSE"] - ref_mse) <= 1e-6 * ref_mse and abs(m["WAPE"] - ref_wape) <= 1e-6 * ref_wape
    with pytest.raises(RuntimeError):
        series_metrics(a, b)                                                        # CPU tensors: no fallback


def test_run_inference_writes_reference_formats(tmp_path):
    """infer.py:66-123 on the fused sampler: two batches, saved arrays have the reference's shapes / dtypes, the
    generated series equal a direct sampler call with the same generator state, metrics equal the oracle's."""
    (dit, _), (vae, _) = make_dit(31), make_vae(32)
    smp = T2SSampler(dit, vae)
    L, steps = 48, 4
    gen = torch.Generator().manual_seed(9)
    batches = [("text", torch.rand(3, L, generator=gen), synth.make_text_embeddings(3, seed=40)),
               ("text", torch.rand(2, L, generator=gen), synth.make_text_embeddings(2, seed=41))]
    g1 = torch.Generator(device=DEV).manual_seed(123)
    x_1, x_t, dec, enc, metrics = run_inference(smp, vae, batches, "flowmatching", steps, 7.0, save_path=str(tmp_path / "run_0"), generator=g1)
    assert x_1.shape == (5, L, 1) and x_t.shape == (5, L, 1) and dec.shape == (5, 64, 30) and enc.shape == (5, 64, 30)
    assert x_1.dtype == np.float32 and x_t.dtype == np.float32
    saved = load_generation(str(tmp_path / "run_0"))
    assert sorted(os.listdir(tmp_path / "run_0")) == ["x_1.npy", "x_t.npy", "x_t_latent_dec_array.npy", "x_t_latent_enc_array.npy"]
    assert np.array_equal(saved["x_t"], x_t) and np.array_equal(saved["x_1"], x_1)
    g2 = torch.Generator(device=DEV).manual_seed(123)
    direct = torch.cat([smp.sample(b[2].to(DEV), L, steps=steps, cfg_scale=7.0, generator=g2) for b in batches], 0)
    assert np.array_equal(direct.cpu().numpy(), x_t[:, :, 0])
    o, x = np.transpose(x_1, (0, 2, 1)), np.transpose(x_t, (0, 2, 1))               # evaluation.py:295-296
    assert abs(metrics["MSE"] - O.calculate_mse(o, x)) <= 1e-6 * O.calculate_mse(o, x)
    assert abs(metrics["WAPE"] - O.calculate_wape(o, x)) <= 1e-6 * O.calculate_wape(o, x)


def test_backbone_classes_on_cuda_match_reference_golden():
    """RectifiedFlow.euler / create_flow and DDPM.q_sample / p_sample (model/backbone/*.py) on CUDA tensors, through the custom
    ops and the C-ABI kernels, against outputs of the reference classes (tests/golden/backbone.npz).  p_sample draws its
    noise with torch.randn inside, like the reference: the golden draw is substituted to compare values."""
    import unittest.mock as um
    from conftest import T, load_golden
    from gpu_util import DEV, max_abs
    from oracle import t2s_oracle as O
    from t2ms_b200 import DDPM, RectifiedFlow
    g = load_golden("backbone.npz")
    x1, t, eps, ti = (T(g[k]).to(DEV) for k in ("x1", "t", "eps", "ti"))
    rf, dd = RectifiedFlow(), DDPM(1000, DEV)
    assert max_abs(rf.euler(x1, eps, 0.01), T(g["euler"])) == 0.0
    x_t, x_0 = rf.create_flow(x1, t)
    assert x_0.shape == x1.shape and abs(float(x_0.mean())) < 0.05 and abs(float(x_0.std()) - 1) < 0.05
    assert max_abs(x_t, O.rf_create_flow(x1.cpu(), t.cpu(), x_0.cpu())) < 1e-6
    with um.patch.object(torch, "randn_like", lambda x: T(g["x0"]).to(DEV)):
        x_t, x_0 = rf.create_flow(x1, t)
    assert max_abs(x_0, T(g["x0"])) == 0.0 and max_abs(x_t, T(g["x_t"])) < 1e-6
    q, e = dd.q_sample(x1, ti, eps)
    # the coefficients sqrt(alpha_bar_t), sqrt(1 - alpha_bar_t) come from torch's CUDA pow here and its CPU pow in the golden run
    assert e is eps and max_abs(q, T(g["q_sample"])) < 1e-5
    mean, var = dd.q_xt_x0(x1, ti)
    assert max_abs(mean + var ** 0.5 * eps, T(g["q_sample"])) < 1e-5
    with um.patch.object(torch, "randn", lambda *a, **k: T(g["p_noise"]).to(DEV)):
        p = dd.p_sample(x1, eps, ti)
    assert max_abs(p, T(g["p_sample"])) < 1e-5
    p2 = dd.p_sample(x1, eps, ti)                                  # its own draw: another sample of the same distribution
    assert max_abs(p2, p) > 1e-3 and torch.isfinite(p2).all()
    for bb in ("beta", "alpha", "alpha_bar"):
        assert max_abs(getattr(dd, bb), T(g[bb])) < 1e-6              # CUDA cumprod's summation order vs the CPU's


def test_custom_ops_trace_under_torch_compile():
    """The module forwards go through torch.library custom ops with fake implementations, so a function that calls them can be
    traced by torch.compile (fullgraph, the aot_eager backend: Dynamo + AOTAutograd with fake tensors, no code generation) and
    returns what the eager call returns."""
    from gpu_util import DEV, make_dit, make_vae, max_abs
    from t2ms_b200 import synth
    (dit, _), (vae, _) = make_dit(3), make_vae(4)
    x, emb = synth.make_noise(4, seed=1).to(DEV), synth.make_text_embeddings(4, seed=2).to(DEV)
    t = torch.tensor([0.1, 0.4, 0.7, 0.9], device=DEV)

    ws, h_dit, h_dec = dit.workspace(4, x.device), dit.packed().handle, vae.decoder._packed_weights().handle   # host-side state: outside the graph

    def step(x, t, emb):
        u = torch.ops.t2s_b200.dit_forward(x, t * 100.0, None, ws, h_dit, 30)
        c = torch.ops.t2s_b200.dit_forward(x, t * 100.0, emb, ws, h_dit, 30)
        x1 = torch.ops.t2s_b200.rf_euler(x, u + 7.0 * (c - u), 0.01)
        return torch.ops.t2s_b200.vae_decode(x1, 48, h_dec)[0]

    with torch.no_grad():
        eager = step(x, t, emb)
        compiled = torch.compile(step, backend="aot_eager", fullgraph=True)(x, t, emb)
    assert eager.shape == (4, 48) and max_abs(eager, compiled) == 0.0
