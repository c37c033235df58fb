"""The CPU oracle (oracle/t2s_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py)."""
import math
import sys

import numpy as np
import pytest
import torch

from conftest import T, load_golden
from oracle import t2s_oracle as O
from t2ms_b200 import synth


def close(a, b, atol, rtol=0.0):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item()
    assert err <= atol + rtol * b.abs().max().item(), f"max-abs err {err}"


def test_weights_reproduce():
    g = load_golden("dit_forward.npz")
    sd = synth.make_dit_state(int(g["dit_seed"]), bias_std=float(g["bias_std"]))
    assert synth.state_checksum(sd) == str(g["dit_checksum"])
    v = load_golden("vae.npz")
    assert synth.state_checksum(synth.make_vae_state(int(v["vae_seed"]))) == str(v["vae_checksum"])
    assert len(sd) == 55
    assert torch.equal(O.sinusoidal_pos_embed(480, 128), sd["pos_embed"])


def test_time_embedding():
    g = load_golden("dit_forward.npz")
    close(O.time_embedding(T(g["t_float"])), g["temb_float"], 1e-6)
    close(O.time_embedding(T(g["t_int"])), g["temb_int"], 1e-6)


def test_dit_forward():
    g = load_golden("dit_forward.npz")
    sd = synth.make_dit_state(int(g["dit_seed"]), bias_std=float(g["bias_std"]))
    x, emb = T(g["x"]), T(g["emb"])
    with torch.no_grad():
        out, hid = O.dit_forward(sd, x, T(g["t_float"]), emb, return_hidden=True)
        close(out, g["cond_float"], 2e-5)
        close(O.dit_forward(sd, x, T(g["t_float"]), None), g["uncond_float"], 2e-5)
        close(O.dit_forward(sd, x, T(g["t_int"]), emb), g["cond_int"], 2e-5)
        tok = g["hidden_tok"].tolist()
        for l in range(4):
            close(hid[l + 1][:, tok, :], g["hidden"][l], 2e-5)


@pytest.mark.parametrize("L", [24, 48, 96])
def test_vae(L):
    g = load_golden("vae.npz")
    sd = synth.make_vae_state(int(g["vae_seed"]))
    with torch.no_grad():
        z, before = O.vae_encode(sd, T(g[f"series_{L}"]))
        close(z, g[f"z_{L}"], 1e-5)
        close(before, g[f"before_{L}"], 1e-5)
        rec, after = O.vae_decode(sd, T(g[f"z_{L}"]), L)
        close(rec, g[f"rec_{L}"], 1e-5)
        close(after, g[f"after_{L}"], 1e-5)
        close(O.vae_decode(sd, T(g["zlat"]), L)[0], g[f"dec_noise_{L}"], 1e-5)
    if L == 48:
        r1 = O.vae_decode(sd, T(g["zlat"])[:1], 48)[0]
        assert r1.shape == (48,)          # torch.squeeze drops the batch dim at B == 1 (vqvae.py:105)
        close(r1, g["dec_b1_48"], 1e-5)


def test_backbone():
    g = load_golden("backbone.npz")
    x1, t, x0, eps, ti = (T(g[k]) for k in ("x1", "t", "x0", "eps", "ti"))
    close(O.rf_create_flow(x1, t, x0), g["x_t"], 1e-7)
    close(O.rf_euler(x1, eps, 0.01), g["euler"], 1e-7)
    close(O.mse(x1, eps), g["rf_loss"], 1e-6)
    sched = O.ddpm_schedule(1000)
    for a, k in zip(sched, ("beta", "alpha", "alpha_bar")):
        assert np.array_equal(a.numpy(), g[k])
    close(O.ddpm_q_sample(x1, ti, eps, sched), g["q_sample"], 1e-6)
    close(O.ddpm_p_sample(x1, eps, ti, T(g["p_noise"]), sched), g["p_sample"], 1e-5)
    assert abs(sched[2][-1].item() - 4.04e-5) < 1e-6


def test_sampling_loops():
    g = load_golden("sampling.npz")
    dit = synth.make_dit_state(int(g["dit_seed"]), bias_std=float(g["bias_std"]))
    vae = synth.make_vae_state(int(g["vae_seed"]))
    assert synth.state_checksum(dit) == str(g["dit_checksum"])
    emb = T(g["emb"])
    for steps, L in ((4, 24), (6, 48), (5, 96)):
        k = f"rf_{steps}_{L}_"
        lat, ser, vel = O.rf_sample(dit, vae, T(g[k + "noise"]), emb, steps, float(g[k + "cfg"]), L,
                                    return_velocities=True)
        close(torch.stack(vel), g[k + "vel"], 2e-4)
        close(lat, g[k + "latent"], 2e-4)
        close(ser, g[k + "series"], 2e-4)
    steps, L = int(g["ddpm_steps"]), int(g["ddpm_L"])
    lat, ser, eps = O.ddpm_sample(dit, vae, T(g["ddpm_noise"]), emb, steps, float(g["ddpm_cfg"]),
                                  T(g["ddpm_step_noise"]), L, return_eps=True)
    close(torch.stack(eps), g["ddpm_eps"], 1e-6, rtol=1e-4)
    close(lat, g["ddpm_latent"], 1e-6, rtol=1e-4)
    close(ser, g["ddpm_series"], 1e-6, rtol=1e-4)


def test_rf_timesteps_match_infer_formula():
    tt = O.rf_timesteps(100, 2)
    for j in (0, 1, 33, 99):
        ref = torch.round(torch.full((2,), j * 1.0 / 100) * 100) / 100
        assert torch.equal(tt[j], ref)


def test_train_step():
    g = load_golden("train.npz")
    dit = synth.make_dit_state(int(g["dit_seed"]), bias_std=float(g["bias_std"]))
    vae = synth.make_vae_state(int(g["vae_seed"]))
    assert synth.state_checksum(dit) == str(g["dit_checksum"])
    series, emb, x0, t = (T(g[k]) for k in ("series", "emb", "x0", "t"))
    with torch.no_grad():
        x1, _ = O.vae_encode(vae, series)
    close(x1, g["x1"], 1e-5)
    x_t = O.rf_create_flow(x1, t, x0)
    loss, grads = O.train_step_grads(dit, x_t, t, emb, x1 - x0)
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * max(1.0, abs(float(g["loss"])))
    names = [str(n)[len("grad_norm/"):] for n in g["names"]]
    assert sorted(names) == sorted(grads.keys())          # 48 tensors receive a gradient
    assert sum(v.numel() for v in grads.values()) == 925592
    for n, ref in zip(names, g["norms"]):
        assert abs(grads[n].norm().item() - ref) <= 1e-4 * max(ref, 1e-3), n
    for k in g.files:
        if k.startswith("grad/"):
            n = k[len("grad/"):]
            close(grads[n].reshape(-1)[:256], g[k], 1e-7, rtol=2e-4)
            p, m, v = O.adamw_step(dit[n], grads[n], torch.zeros_like(dit[n]), torch.zeros_like(dit[n]), 1, 1e-4)
            close(p.reshape(-1)[:256], g["param_after/" + n], 2e-7)


def test_series_metrics_oracle_matches_reference_functions():
    """oracle calculate_mse / calculate_wape against the reference's own functions (evaluation.py:166-206) on the
    transposed (N, 1, L) layout evaluation.py:295-296 feeds them, incl. a zero-denominator sample (nanmean).  The
    reference accumulates in the arrays' float32, the oracle in float64: relative 1e-6."""
    g = load_golden("eval.npz")
    for c in "abc":
        o, x = np.transpose(g[f"{c}/ori"], (0, 2, 1)), np.transpose(g[f"{c}/gen"], (0, 2, 1))
        assert abs(O.calculate_mse(o, x) - float(g[f"{c}/mse"])) <= 1e-6 * float(g[f"{c}/mse"])
        assert abs(O.calculate_wape(o, x) - float(g[f"{c}/wape"])) <= 1e-6 * float(g[f"{c}/wape"])


@pytest.mark.parametrize("dim", [50, 64])
def test_variable_width_dit_oracle_matches_fork_reference(dim):
    """oracle dit_forward on (B,64,dim) latents against the fork's model/denoiser/mytransformer.py Transformer(dim)."""
    g = load_golden("dit_tokens.npz")
    k = f"h{dim}/"
    sd = synth.make_dit_state(40 + dim, bias_std=0.05, dim=dim)
    assert synth.state_checksum(sd) == str(g[k + "checksum"])
    x, emb = synth.make_noise(2, seed=50 + dim, dim=dim), synth.make_text_embeddings(2, seed=60 + dim)
    t_f, t_i = torch.tensor([0.25, 0.9]), torch.tensor([3, 871], dtype=torch.long)
    with torch.no_grad():
        close(O.dit_forward(sd, x, t_f, emb), g[k + "cond_float"], 2e-5)
        close(O.dit_forward(sd, x, t_f, None), g[k + "uncond_float"], 2e-5)
        close(O.dit_forward(sd, x, t_i, emb), g[k + "cond_int"], 2e-5)
        lat, _, vel = O.rf_sample(sd, None, x, emb, 3, 7.0, return_velocities=True)
        close(torch.stack(vel), g[k + "rf_vel"], 1e-4)
        close(lat, g[k + "rf_final"], 1e-4)
        sn = synth.make_step_noise(3, 2, seed=70 + dim, dim=dim)
        lat, _, eps = O.ddpm_sample(sd, None, x, emb, 3, 7.0, sn, return_eps=True)
        close(torch.stack(eps), g[k + "ddpm_eps"], 1e-3, rtol=1e-5)
        close(lat, g[k + "ddpm_final"], 1e-3, rtol=1e-5)


@pytest.mark.parametrize("case,cin,flow", [("uni48", 1, 30), ("multi100", 7, 50), ("multi90", 7, 50)])
def test_lavae_training_oracle_matches_reference(case, cin, flow):
    """oracle lavae_train_grads (loss, recon, z, every parameter gradient, AdamW update) against
    vqvae.shared_eval(batch, optimizer, 'train') of the reference (vqvae.py:118-135) and of the fork (myvqvae.py:116-136)."""
    g = load_golden("vae_train.npz")
    k = case + "/"
    batch = T(g[k + "batch"])
    sd = synth.make_vae_state(80 + batch.shape[-1], in_channels=cin)
    assert synth.state_checksum(sd) == str(g[k + "checksum"])
    loss, recon_error, recon, z, grads = O.lavae_train_grads(sd, batch, flow_dim=flow)
    assert abs(float(loss) - float(g[k + "loss"])) <= 1e-6 and abs(float(recon_error) - float(g[k + "recon_error"])) <= 1e-6
    close(recon, g[k + "recon"], 1e-5)
    close(z, g[k + "z"], 1e-5)
    names = [str(n) for n in g[k + "names"]]
    assert sorted(names) == sorted(sd.keys())
    for n, ref_norm in zip(names, g[k + "grad_norms"]):
        assert abs(float(grads[n].norm()) - float(ref_norm)) <= 1e-4 * max(1.0, float(ref_norm)), n
        close(grads[n].reshape(-1)[:128], g[k + "grad/" + n], 1e-5, rtol=1e-4)
        # core.py:15: AdamW(lr 1e-3, weight_decay 1e-2), first step (LinearLR start_factor 0.1 -> lr 1e-4)
        p, _, _ = O.adamw_step(sd[n].clone(), grads[n], torch.zeros_like(sd[n]), torch.zeros_like(sd[n]), 1, 1e-4, wd=1e-2)
        close(p.reshape(-1)[:128], g[k + "after/" + n], 2e-6)


@pytest.mark.parametrize("dim", [50, 64])
def test_variable_width_training_oracle_matches_fork_reference(dim):
    """oracle train_step_grads on (B,64,dim) latents against one rectified-flow training step of the fork's
    Transformer(dim) (mytrain.py:66-87, model/denoiser/mytransformer.py)."""
    g = load_golden("dit_tokens_train.npz")
    k = f"h{dim}/"
    sd = synth.make_dit_state(140 + dim, bias_std=0.02, dim=dim)
    assert synth.state_checksum(sd) == str(g[k + "checksum"])
    x1, x0 = synth.make_noise(3, seed=150 + dim, dim=dim), synth.make_noise(3, seed=160 + dim, dim=dim)
    emb, t = synth.make_text_embeddings(3, seed=170 + dim), torch.tensor([0.2, 0.55, 0.9])
    x_t, target = O.rf_create_flow(x1, t, x0), x1 - x0
    loss, grads = O.train_step_grads(sd, x_t, t, emb, target)
    assert abs(float(loss) - float(g[k + "loss"])) <= 1e-5 * float(g[k + "loss"])
    for n, ref_norm in zip([str(x) for x in g[k + "names"]], g[k + "grad_norms"]):
        assert abs(float(grads[n].norm()) - float(ref_norm)) <= 2e-4 * max(float(ref_norm), 1e-8), n
        close(grads[n].reshape(-1)[:128], g[k + "grad/" + n], 1e-7, rtol=2e-4)


def test_philox_known_answer_vectors():
    """Philox4x32-10 against the published Random123 known-answer vectors (kat_vectors: philox4x32 10): the generator behind
    the in-kernel DDPM noise (t2s_sample_ddpm_seeded) and its numpy restatement."""
    f = 0xFFFFFFFF
    for ctr, key, expect in (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                             ((f, f, f, f), (f, f), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))):
        out = O.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], *key)
        assert tuple(int(w[0]) for w in out) == expect
    z = O.philox_normal(987654321, 7, 400_000)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01 and np.isfinite(z).all()
    assert not np.array_equal(z[:1000], O.philox_normal(987654321, 8, 1000))           # the step is part of the counter


def test_staged_reference_modules_agree_with_the_oracle():
    """baseline/_ref (the unmodified reference files staged by oracle/ref_install.py for the bench baselines) drives the
    same loop as the oracle: equal to fp32 rounding.  Skipped where the reference was never mounted."""
    from oracle import ref_install as R
    if not R.available():
        pytest.skip("baseline/_ref is not staged")
    from t2ms_b200 import synth
    ref = R.load()
    dsd, vsd = synth.make_dit_state(0), synth.make_vae_state(1)
    dit, vae = R.build_reference_models(ref, dsd, vsd)
    emb, noise = synth.make_text_embeddings(3), synth.make_noise(3)
    s = R.reference_sample(ref, dit, vae, emb, noise, 3, 7.0, 48)
    _, so = O.rf_sample(dsd, vsd, noise, emb, 3, 7.0, 48)
    assert (s - so).abs().max().item() < 1e-5
    sn = synth.make_step_noise(3, 3)
    s = R.reference_sample(ref, dit, vae, emb, noise, 3, 7.0, 24, backbone="ddpm", step_noise=sn)
    _, so = O.ddpm_sample(dsd, vsd, noise, emb, 3, 7.0, sn, 24)
    assert (s - so).abs().max().item() < 1e-5
    for m in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[m]                       # leave no reference modules behind for the compat-alias tests
