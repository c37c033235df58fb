"""GPU parity of the fused sampling loops + LA-VAE against the oracle and the reference golden vectors.

Tolerances (BASELINE.json north_star): per-step guided velocity / epsilon within 2e-3 relative (L2),
final generated series within 1e-2 max-abs; LA-VAE (fp32) within 1e-5 relative to the output scale.
"""
import pytest
import torch

from conftest import T, load_golden
from oracle import t2s_oracle as O

pytestmark = pytest.mark.gpu

TOL_STEP = 2e-3
TOL_SERIES = 1e-2


@pytest.mark.parametrize("L", [24, 48, 96])
def test_vae_matches_reference_golden(L):
    from gpu_util import DEV, make_vae, max_abs
    g = load_golden("vae.npz")
    vae, sd = make_vae(int(g["vae_seed"]))
    with torch.no_grad():
        z, before = vae.encoder(T(g[f"series_{L}"]).to(DEV))
        assert z.shape == (3, 64, 30) and before.shape == (3, 64, L // 4)
        assert max_abs(z, T(g[f"z_{L}"])) < 1e-5 and max_abs(before, T(g[f"before_{L}"])) < 1e-5
        rec, after = vae.decoder(T(g[f"z_{L}"]).to(DEV), length=L)
        assert rec.shape == (3, L)
        assert max_abs(rec, T(g[f"rec_{L}"])) < 1e-5 and max_abs(after, T(g[f"after_{L}"])) < 1e-5
        rec2, _ = vae.decoder(T(g["zlat"]).to(DEV), length=L)
        assert max_abs(rec2, T(g[f"dec_noise_{L}"])) < 2e-5
        if L == 48:
            r1, _ = vae.decoder(T(g["zlat"])[:1].to(DEV), length=48)
            assert r1.shape == (48,)                       # torch.squeeze at B == 1 (vqvae.py:105)
            assert max_abs(r1, T(g["dec_b1_48"])) < 2e-5


def test_vae_large_batch_vs_oracle():
    from gpu_util import DEV, make_vae, max_abs
    from t2ms_b200 import synth
    vae, sd = make_vae(7)
    z = synth.make_noise(257, seed=8)
    with torch.no_grad():
        for L in (24, 96):
            ref, ref_after = O.vae_decode(sd, z, L)
            out, after = vae.decoder(z.to(DEV), length=L)
            assert max_abs(out, ref) < 2e-5 and max_abs(after, ref_after) < 1e-5
        s = synth.make_series(130, 96, seed=9)
        zr, br = O.vae_encode(sd, s)
        zo, bo = vae.encoder(s.to(DEV))
        assert max_abs(zo, zr) < 1e-5 and max_abs(bo, br) < 1e-5


def test_rf_sampling_matches_reference_golden():
    from gpu_util import DEV, make_dit, make_vae, max_abs, rel_l2
    from t2ms_b200 import T2SSampler
    g = load_golden("sampling.npz")
    dit, _ = make_dit(int(g["dit_seed"]), bias_std=float(g["bias_std"]))
    vae, _ = make_vae(int(g["vae_seed"]))
    smp = T2SSampler(dit, vae)
    emb = T(g["emb"]).to(DEV)
    for steps, L in ((4, 24), (6, 48), (5, 96)):
        k = f"rf_{steps}_{L}_"
        noise = T(g[k + "noise"]).to(DEV)
        lat, trace = smp.sample_latent(emb, steps=steps, cfg_scale=float(g[k + "cfg"]), noise=noise, trace=True)
        ser = smp.sample(emb, L, steps=steps, cfg_scale=float(g[k + "cfg"]), noise=noise)
        vel = T(g[k + "vel"])
        per_step = [rel_l2(trace[j], vel[j]) for j in range(steps)]
        print(k, "per-step rel-L2", ["%.2e" % e for e in per_step], "series max-abs %.2e" % max_abs(ser, T(g[k + "series"])))
        assert max(per_step) < TOL_STEP
        assert max_abs(lat, T(g[k + "latent"])) < TOL_SERIES
        assert ser.shape == (2, L) and max_abs(ser, T(g[k + "series"])) < TOL_SERIES
        assert torch.equal(noise, T(g[k + "noise"]).to(DEV))          # the caller's noise tensor is not modified


def test_ddpm_sampling_matches_reference_golden():
    from gpu_util import DEV, make_dit, make_vae, max_abs, rel_l2
    from t2ms_b200 import T2SSampler
    g = load_golden("sampling.npz")
    dit, _ = make_dit(int(g["dit_seed"]), bias_std=float(g["bias_std"]))
    vae, _ = make_vae(int(g["vae_seed"]))
    smp = T2SSampler(dit, vae)
    steps, L, cfg = int(g["ddpm_steps"]), int(g["ddpm_L"]), float(g["ddpm_cfg"])
    emb = T(g["emb"]).to(DEV)
    lat, trace = smp.sample_latent(emb, steps=steps, cfg_scale=cfg, backbone="ddpm", noise=T(g["ddpm_noise"]).to(DEV),
                                   step_noise=T(g["ddpm_step_noise"]).to(DEV), trace=True)
    eps = T(g["ddpm_eps"])
    per_step = [rel_l2(trace[j], eps[j]) for j in range(steps)]
    print("ddpm per-step rel-L2", ["%.2e" % e for e in per_step])
    assert max(per_step) < TOL_STEP
    ref_lat = T(g["ddpm_latent"])
    assert rel_l2(lat, ref_lat) < TOL_STEP
    ser = smp.sample(emb, L, steps=steps, cfg_scale=cfg, backbone="ddpm", noise=T(g["ddpm_noise"]).to(DEV),
                     step_noise=T(g["ddpm_step_noise"]).to(DEV))
    scale = max(1.0, T(g["ddpm_series"]).abs().max().item())
    assert max_abs(ser, T(g["ddpm_series"])) < TOL_SERIES * scale


def test_rf_sampling_vs_oracle_config1():
    """BASELINE config 1: RF, length 24, batch 8 (reference CPU case), 10 steps for test time."""
    from gpu_util import DEV, make_dit, make_vae, max_abs, rel_l2
    from t2ms_b200 import T2SSampler, synth
    dit, dsd = make_dit(51)
    vae, vsd = make_vae(52)
    B, steps = 8, 10
    emb, noise = synth.make_text_embeddings(B, 53), synth.make_noise(B, 54)
    lat_ref, ser_ref, vel_ref = O.rf_sample(dsd, vsd,
                                            noise, emb, steps, 7.0, 24, return_velocities=True)
    smp = T2SSampler(dit, vae)
    lat, trace = smp.sample_latent(emb.to(DEV), steps=steps, cfg_scale=7.0, noise=noise.to(DEV), trace=True)
    ser = smp.sample(emb.to(DEV), 24, steps=steps, cfg_scale=7.0, noise=noise.to(DEV))
    per_step = [rel_l2(trace[j], vel_ref[j]) for j in range(steps)]
    print("config1 per-step rel-L2", ["%.2e" % e for e in per_step], "series max-abs %.2e" % max_abs(ser, ser_ref))
    assert max(per_step) < TOL_STEP
    assert max_abs(ser, ser_ref) < TOL_SERIES


def test_full_size_properties():
    """BASELINE config 2 size (B=1024, L=96) through size-independent properties: samples are independent,
    so any chunking reproduces the same series bit-for-bit; duplicates of one prompt+noise produce identical rows;
    a sub-batch small enough for the latency kernels (attention with 96-key instead of 48-key chunks: another summation
    order) agrees to the attention kernel's own accuracy; and a small slice matches the oracle."""
    from gpu_util import DEV, make_dit, make_vae, max_abs
    from t2ms_b200 import T2SSampler, synth
    dit, dsd = make_dit(61)
    vae, vsd = make_vae(62)
    B, steps = 1024, 3
    emb, noise = synth.make_text_embeddings(B, 63).to(DEV), synth.make_noise(B, 64).to(DEV)
    emb[1000:] = emb[0]
    noise[1000:] = noise[0]
    smp = T2SSampler(dit, vae)
    full = smp.sample(emb, 96, steps=steps, noise=noise)
    assert full.shape == (B, 96) and torch.isfinite(full).all()
    assert all(torch.equal(full[i], full[0]) for i in range(1000, B))
    chunked = smp.sample(emb, 96, steps=steps, noise=noise, chunk=37)
    assert torch.equal(full, chunked)
    sub = smp.sample(emb[500:507], 96, steps=steps, noise=noise[500:507])
    assert max_abs(full[500:507], sub) < TOL_SERIES / 5
    sub = smp.sample(emb[500:580], 96, steps=steps, noise=noise[500:580])            # large enough for the throughput kernels
    assert torch.equal(full[500:580], sub)
    _, ser_ref = O.rf_sample(dsd, vsd, noise[:4].cpu(), emb[:4].cpu(), steps, 7.0, 96)
    assert max_abs(full[:4], ser_ref) < TOL_SERIES


def test_ddpm_step_noise_windows():
    """Without caller-supplied step noise the sampler draws it per window of steps and enqueues the loop window by window
    (t100 / coef offset into the same C entry): equal to one call with the concatenated noise of the same generator."""
    from gpu_util import DEV, make_dit, max_abs
    from t2ms_b200 import T2SSampler, synth
    dit, _ = make_dit(3)
    smp = T2SSampler(dit)
    B, steps = 3, 7
    emb, x0 = synth.make_text_embeddings(B, seed=5).to(DEV), synth.make_noise(B, seed=6).to(DEV)
    smp.NOISE_WINDOW_BYTES = 3 * B * 1920 * 4                      # windows of 3, 3, 1 steps
    g = torch.Generator(device=DEV).manual_seed(11)
    lat, tr = smp.sample_latent(emb, steps=steps, backbone="ddpm", noise=x0, generator=g, trace=True, noise_source="torch")
    g = torch.Generator(device=DEV).manual_seed(11)
    sn = torch.cat([torch.randn(n, B, 64, 30, device=DEV, generator=g) for n in (3, 3, 1)])
    smp.NOISE_WINDOW_BYTES = 1 << 30
    lat2, tr2 = smp.sample_latent(emb, steps=steps, backbone="ddpm", noise=x0, step_noise=sn, trace=True)
    assert max_abs(lat, lat2) == 0.0 and max_abs(tr, tr2) == 0.0


def test_sampling_loop_is_cuda_graph_capturable():
    """t2s_sample only enqueues kernels on the caller's stream (no allocation, no synchronisation): the whole guided loop
    + decode can be captured once in a CUDA graph and replayed; the replay equals the eager call bit for bit."""
    from gpu_util import DEV, make_dit, make_vae, max_abs
    from t2ms_b200 import T2SSampler, synth
    (dit, _), (vae, _) = make_dit(3), make_vae(4)
    smp = T2SSampler(dit, vae)
    B, steps, L = 5, 6, 48
    emb, x0 = synth.make_text_embeddings(B, seed=5).to(DEV), synth.make_noise(B, seed=6).to(DEV)
    eager = smp.sample(emb, L, steps=steps, noise=x0)              # also warms up: packed weights, workspace, tables
    torch.cuda.synchronize()
    static_emb, static_x0 = emb.clone(), x0.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        smp.sample(static_emb, L, steps=steps, noise=static_x0)    # allocator warm-up on the capture stream
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        out = smp.sample(static_emb, L, steps=steps, noise=static_x0)
    g.replay()
    torch.cuda.synchronize()
    assert max_abs(out, eager) == 0.0
    static_emb.copy_(synth.make_text_embeddings(B, seed=7).to(DEV))   # new inputs through the static buffers
    g.replay()
    torch.cuda.synchronize()
    ref2 = smp.sample(static_emb, L, steps=steps, noise=static_x0)
    assert max_abs(out, ref2) == 0.0


def test_sample_graph_and_host_entry_match_eager():
    """`sample_graph` (captured graph, static buffers) and the host-buffer entry that uses it for small batches return
    what the plain enqueueing path returns; a weight update invalidates the cached graph."""
    from gpu_util import DEV, make_dit, make_vae, max_abs
    from t2ms_b200 import T2SSampler, synth
    (dit, _), (vae, _) = make_dit(3), make_vae(4)
    smp = T2SSampler(dit, vae)
    B, steps, L = 4, 5, 24
    emb, x0 = synth.make_text_embeddings(B, seed=5), synth.make_noise(B, seed=6).to(DEV)
    eager = smp.sample(emb.to(DEV), L, steps=steps, noise=x0)
    for _ in range(2):                                              # capture, then a pure replay
        assert max_abs(smp.sample_graph(emb.to(DEV), L, steps=steps, noise=x0), eager) == 0.0
    host = smp.sample_host(emb.pin_memory(), L, steps=steps, noise=x0)
    assert not host.is_cuda and max_abs(host, eager) == 0.0
    assert max_abs(smp.sample_host(emb.pin_memory(), L, steps=steps, noise=x0, graph=False), eager) == 0.0
    with torch.no_grad():
        dit.ln.bias.add_(0.1)                                       # new weights -> new packed object -> new graph
    eager2 = smp.sample(emb.to(DEV), L, steps=steps, noise=x0)
    assert max_abs(eager2, eager) > 0 and max_abs(smp.sample_graph(emb.to(DEV), L, steps=steps, noise=x0), eager2) == 0.0


def test_ddpm_in_kernel_philox_noise_matches_its_restatement():
    """DDPM with the ancestral noise generated inside the update kernel (t2s_sample_ddpm_seeded): the whole 6-step guided loop
    equals the CPU oracle fed with the numpy restatement of the same counter-based generator (oracle.philox_normal), the
    noise is standard normal, and a second call with the same seed reproduces the result bit for bit."""
    import numpy as np
    from gpu_util import DEV, make_dit, max_abs, rel_l2
    from t2ms_b200 import T2SSampler, synth
    dit, dsd = make_dit(3)
    smp = T2SSampler(dit)
    B, steps, seed = 3, 6, 0x1234_5678_9ABC_DEF1
    emb, x0 = synth.make_text_embeddings(B, seed=5), synth.make_noise(B, seed=6)
    lat, tr = smp.sample_latent(emb.to(DEV), steps=steps, backbone="ddpm", noise=x0.to(DEV), trace=True, seed=seed)
    lat2 = smp.sample_latent(emb.to(DEV), steps=steps, backbone="ddpm", noise=x0.to(DEV), seed=seed)
    assert max_abs(lat, lat2) == 0.0
    sn = torch.from_numpy(np.stack([O.philox_normal(seed, j, B * 1920) for j in range(steps)])).reshape(steps, B, 64, 30)
    assert abs(sn.mean().item()) < 0.02 and abs(sn.std().item() - 1.0) < 0.02
    ref, _, eps = O.ddpm_sample(dsd, None, x0, emb, steps, 7.0, sn, return_eps=True)
    per = [rel_l2(tr[j], eps[j]) for j in range(steps)]
    assert max(per) < 2e-3, per
    assert rel_l2(lat, ref) < 2e-3
    other = smp.sample_latent(emb.to(DEV), steps=steps, backbone="ddpm", noise=x0.to(DEV), seed=seed + 1)
    assert max_abs(other, lat) > 1e-3                               # another seed, another trajectory
    default = smp.sample_latent(emb.to(DEV), steps=2, backbone="ddpm", noise=x0.to(DEV))   # seed drawn from torch's RNG
    assert torch.isfinite(default).all()


def test_cached_graphs_survive_workspace_churn_and_weight_updates():
    """a captured graph bakes in the workspace and packed-weight pointers.  The cache entry keeps those
    objects alive, keys on pack generations (not id()), and drops graphs of replaced weights: six batch sizes + interleaved
    forwards at other sizes (which evict the module's workspace cache) + an optimizer-style weight update, then every
    cached graph still reproduces the eager result."""
    from gpu_util import DEV, make_dit, make_vae, max_abs
    from t2ms_b200 import T2SSampler, synth
    (dit, _), (vae, _) = make_dit(3), make_vae(4)
    smp = T2SSampler(dit, vae)
    steps, L = 3, 24
    cases = {}
    for B in (1, 2, 3, 5, 6, 7):
        emb, x0 = synth.make_text_embeddings(B, seed=10 + B).to(DEV), synth.make_noise(B, seed=20 + B).to(DEV)
        cases[B] = (emb, x0, smp.sample(emb, L, steps=steps, noise=x0))
        assert max_abs(smp.sample_graph(emb, L, steps=steps, noise=x0), cases[B][2]) == 0.0
        with torch.no_grad():                                       # other sizes: the module's small workspace cache turns over
            dit(input=synth.make_noise(B + 8, seed=1).to(DEV), t=torch.zeros(B + 8, device=DEV), text_input=None)
    assert len(smp._graphs) == 6
    junk = [torch.randn(1 << 20, device=DEV) for _ in range(8)]     # anything freed by mistake would be handed out again here
    for B, (emb, x0, ref) in cases.items():
        assert max_abs(smp.sample_graph(emb, L, steps=steps, noise=x0), ref) == 0.0
    del junk
    with torch.no_grad():
        dit.ln.bias.add_(0.05)                                      # new weights: a new pack generation
    emb, x0, old = cases[3]
    new = smp.sample(emb, L, steps=steps, noise=x0)
    assert max_abs(new, old) > 0 and max_abs(smp.sample_graph(emb, L, steps=steps, noise=x0), new) == 0.0
    assert all(k[5] == dit.packed().generation for k in smp._graphs)   # graphs of the replaced weights are gone
