"""GPU parity of the variable-width denoiser: the fork's ``Transformer(dim)``
(model/denoiser/mytransformer.py:127-204, dim = config.yaml flow_dim 50 / 64 -> 800 / 1024 tokens) through the same
kernels instantiated per latent width, against golden vectors generated from the fork's module and against the oracle.

Tolerances as for the T2S shape (BASELINE.json north_star): single forward rel-L2 <= 2e-3, per-step guided
velocity / epsilon rel-L2 <= 2e-3, final latent max-abs <= 1e-2."""
import pytest
import torch

from conftest import T, load_golden
from gpu_util import DEV, max_abs, rel_l2
from oracle import t2s_oracle as O
from t2ms_b200 import T2SSampler, Transformer, synth

pytestmark = pytest.mark.gpu


def _model(dim, seed, bias_std=0.05):
    sd = synth.make_dit_state(seed, bias_std=bias_std, dim=dim)
    m = Transformer(dim)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


@pytest.mark.parametrize("dim", [50, 64])
def test_forward_matches_fork_golden(dim):
    g = load_golden("dit_tokens.npz")
    k = f"h{dim}/"
    m, _ = _model(dim, 40 + dim)
    assert m.pos_embed.shape == (1, 16 * dim, 128) and len(m.state_dict()) == 55
    x, emb = synth.make_noise(2, seed=50 + dim, dim=dim).to(DEV), synth.make_text_embeddings(2, seed=60 + dim).to(DEV)
    t_f, t_i = torch.tensor([0.25, 0.9], device=DEV), torch.tensor([3, 871], dtype=torch.long, device=DEV)
    with torch.no_grad():
        for out, key in ((m(input=x, t=t_f, text_input=emb), "cond_float"), (m(input=x, t=t_f, text_input=None), "uncond_float"),
                         (m(input=x, t=t_i, text_input=emb), "cond_int")):
            assert out.shape == (2, 64, dim)
            e = rel_l2(out, T(g[k + key]))
            print(dim, key, "rel-L2 %.2e" % e)
            assert e < 2e-3


@pytest.mark.parametrize("dim", [50, 64])
def test_guided_sampling_matches_fork_golden(dim):
    g = load_golden("dit_tokens.npz")
    k = f"h{dim}/"
    m, _ = _model(dim, 40 + dim)
    smp = T2SSampler(m)
    x, emb = synth.make_noise(2, seed=50 + dim, dim=dim).to(DEV), synth.make_text_embeddings(2, seed=60 + dim).to(DEV)
    lat, tr = smp.sample_latent(emb, steps=3, cfg_scale=7.0, noise=x, trace=True)
    per = [rel_l2(tr[j], T(g[k + "rf_vel"])[j]) for j in range(3)]
    print(dim, "rf per-step", ["%.2e" % e for e in per])
    assert max(per) < 2e-3 and max_abs(lat, T(g[k + "rf_final"])) < 1e-2
    sn = synth.make_step_noise(3, 2, seed=70 + dim, dim=dim).to(DEV)
    lat, tr = smp.sample_latent(emb, steps=3, cfg_scale=7.0, backbone="ddpm", noise=x, step_noise=sn, trace=True)
    per = [rel_l2(tr[j], T(g[k + "ddpm_eps"])[j]) for j in range(3)]
    print(dim, "ddpm per-step", ["%.2e" % e for e in per])
    # DDPM with 3 steps amplifies the prediction error by 1/sqrt(alpha_t) ~ 7: the latent bar is relative to its scale
    assert max(per) < 2e-3 and rel_l2(lat, T(g[k + "ddpm_final"])) < 2e-3
    with pytest.raises(RuntimeError):
        smp.sample(emb, 96, steps=2)                       # built without a decoder


@pytest.mark.parametrize("dim,B", [(50, 5), (64, 37)])
def test_forward_odd_batches_vs_oracle(dim, B):
    """Odd sequence counts (a half-filled last pair) and more sequences than one wave of work items."""
    m, sd = _model(dim, 7, bias_std=0.02)
    x, emb = synth.make_noise(B, seed=8, dim=dim), synth.make_text_embeddings(B, seed=9)
    t = torch.linspace(0, 1, B)
    with torch.no_grad():
        ref = O.dit_forward(sd, x, t, emb)
        out = m(input=x.to(DEV), t=t.to(DEV), text_input=emb.to(DEV))
    assert rel_l2(out, ref) < 2e-3
    per_seq = ((out.cpu().double() - ref.double()).flatten(1).norm(dim=1) / ref.double().flatten(1).norm(dim=1)).max().item()
    assert per_seq < 3e-3


def test_unsupported_width_raises():
    with pytest.raises(ValueError):
        Transformer(40)


def test_fork_generation_loop_with_multivariate_lavae():
    """The fork's generation loop (myinfer.py:116-147): Transformer(flow_dim) guided sampling + the multivariate LA-VAE
    decoder, through T2SSampler.sample, against the oracle (rf_sample on (B,64,50) latents + lavae_decode)."""
    from argparse import Namespace
    from t2ms_b200.mylavae import vqvae as vq_multi
    m, sd = _model(50, 21, bias_std=0.02)
    vae = vq_multi(Namespace(block_hidden_size=128, num_residual_layers=2, res_hidden_size=256, embedding_dim=64, flow_dim=50, input_dim=7))
    vsd = synth.make_vae_state(22, in_channels=7)
    vae.load_state_dict(vsd, strict=True)
    vae = vae.to(DEV).eval()
    m.encoder = vae.encoder                                 # myinfer.py:133
    assert len([n for n, _ in m._own_params()]) == 55
    B, L, steps = 3, 100, 4
    x0, emb = synth.make_noise(B, seed=23, dim=50), synth.make_text_embeddings(B, seed=24)
    series, z = T2SSampler(m, vae).sample(emb.to(DEV), L, steps=steps, cfg_scale=7.0, noise=x0.to(DEV), return_latent=True)
    lat_ref, _ = O.rf_sample(sd, None, x0, emb, steps, 7.0)
    ser_ref, _ = O.lavae_decode(vsd, lat_ref, L)
    assert series.shape == (B, 7, L) and max_abs(z, lat_ref) < 1e-2 and max_abs(series, ser_ref) < 1e-2
    with torch.no_grad():
        z_enc, _ = m.encoder(torch.rand(B, 7, L, device=DEV))   # the attached encoder still runs (myinfer.py:117)
    assert z_enc.shape == (B, 64, 50)


@pytest.mark.parametrize("dim", [50, 64])
def test_wide_training_step_matches_fork_golden(dim):
    """Training step of the fork's Transformer(dim) (mytrain.py:66-87): prediction, loss and every parameter gradient
    against golden vectors from model/denoiser/mytransformer.py.  Same tolerances as the T2S shape
    (tests/test_gpu_train.py): prediction rel-L2 2e-3, loss 2e-3 relative, gradients rel-L2 1e-2 (tf32 / fp16 operands)."""
    from t2ms_b200.training import DitTrainer, trainable_names
    g = load_golden("dit_tokens_train.npz")
    k = f"h{dim}/"
    sd = synth.make_dit_state(140 + dim, bias_std=0.02, dim=dim)
    m = Transformer(dim)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).train()
    x1, x0 = synth.make_noise(3, seed=150 + dim, dim=dim), synth.make_noise(3, seed=160 + dim, dim=dim)
    emb, t = synth.make_text_embeddings(3, seed=170 + dim), torch.tensor([0.2, 0.55, 0.9])
    tr = DitTrainer(m)
    x_t, target = tr.make_inputs("flowmatching", x1.to(DEV), x0.to(DEV), t.to(DEV))
    assert max_abs(x_t, O.rf_create_flow(x1, t, x0)) < 1e-6 and max_abs(target, x1 - x0) < 1e-6
    tr.zero_grad()
    pred = torch.empty(3, 64, dim, device=DEV)
    tr.forward_backward(x_t, t.to(DEV), emb.to(DEV), target, pred=pred)
    torch.cuda.synchronize()
    assert rel_l2(pred, T(g[k + "pred"])) < 2e-3
    loss = tr.loss_sum.item() / (3 * 64 * dim)
    assert abs(loss - float(g[k + "loss"])) <= 2e-3 * float(g[k + "loss"])
    worst = 0.0
    for n, ref_norm in zip([str(x) for x in g[k + "names"]], g[k + "grad_norms"]):
        gr = tr.grads.view(n)
        assert abs(gr.norm().item() - float(ref_norm)) <= 1e-2 * float(ref_norm), (n, gr.norm().item(), float(ref_norm))
        e = rel_l2(gr.reshape(-1)[:128], T(g[k + "grad/" + n]))
        worst = max(worst, e)
        assert e < 1e-2, (n, e)
    assert sorted(str(x) for x in g[k + "names"]) == sorted(trainable_names())
    print(dim, "worst gradient-slice rel-L2 %.2e" % worst)


def test_wide_training_through_autograd_and_optimizer():
    """The reference-style loop of mytrain.py (model(...), loss.backward(), optimizer.step()) on Transformer(64) in training
    mode: gradients equal the oracle's, a step changes the parameters."""
    dim, B = 64, 5
    sd = synth.make_dit_state(9, bias_std=0.02, dim=dim)
    m = Transformer(dim)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).train()
    x1, x0 = synth.make_noise(B, seed=1, dim=dim), synth.make_noise(B, seed=2, dim=dim)
    emb, t = synth.make_text_embeddings(B, seed=3), torch.linspace(0.05, 0.95, B)
    x_t, target = O.rf_create_flow(x1, t, x0), x1 - x0
    opt = torch.optim.AdamW([p for n, p in m.named_parameters() if p.requires_grad], lr=1e-4, weight_decay=0)
    pred = m(input=x_t.to(DEV), t=t.to(DEV), text_input=emb.to(DEV))
    loss = torch.nn.functional.mse_loss(pred, target.to(DEV))
    opt.zero_grad()
    loss.backward()
    loss_ref, grads = O.train_step_grads(sd, x_t, t, emb, target)
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * loss_ref.item()
    named = dict(m.named_parameters())
    for n in ("layers.0.attn.qkv.weight", "layers.3.mlp.fc2.weight", "patch_emb.weight", "layers.2.adaLN_modulation.1.bias", "ln.weight"):
        assert rel_l2(named[n].grad, grads[n]) < 1e-2, n
    before = named["layers.1.attn.proj.weight"].detach().clone()
    opt.step()
    assert (named["layers.1.attn.proj.weight"] - before).abs().max().item() > 0
