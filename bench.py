#!/usr/bin/env python
"""Benchmark of the T2S generation hot path (BASELINE.json metric: sampled series/sec, length 96,
rectified flow, classifier-free guidance, LA-VAE decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: `--batch` series generated
from text embeddings with `--rf-steps` guided denoising steps + decode (BASELINE config 2: batch 1024,
length 96, 100 steps, cfg 7).  N > 1 (torchrun, one rank per GPU): every rank generates its own
`--batch` series (weak scaling) and the step ends with the final all_gather of the series.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner when the box exports NCCL_DEBUG, torchrun's
# OMP notice, ...) write to file descriptor 1 behind Python's back, so the real stdout is kept aside and fd 1 is pointed at
# stderr for the lifetime of the process; emit() is the only writer of the real stdout.
_REAL_STDOUT = None


def capture_stdout() -> None:
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


BASELINE_WHAT = {"reference": "unmodified reference modules from baseline/_ref (torch fp32 CPU, timm Attention -> SDPA)",
                 "port": "oracle port (torch fp32 CPU); baseline/_ref not staged"}
METRIC = "t2s_dit_rf_sampled_series_per_sec"
UNIT = "series/s"
FLOP_FWD = 976_960_512                 # per DiT forward per sequence
FLOP_ATTN_LAYER = 2 * 58_982_400       # QK^T + PV per sequence per block
FLOP_TOKEN_MID = 15_728_640 + 31_457_280 + 31_457_280 + 47_185_920   # proj + fc1 + fc2 + next-block QKV
FLOP_DECODE_96 = 15_360_000
# algorithmic HBM bytes per sequence and launch: attention reads q|k|v images (4 heads x 94,208 B) and writes the
# fp16 attention output (480 x 128 x 2); the MID token kernel reads the residual (fp32) and the attention output,
# writes the residual and the next block's q|k|v images
ALGO_BYTES_PER_SEQ = {"attention": 4 * 94_208 + 480 * 128 * 2, "token_mid": 2 * 480 * 128 * 4 + 480 * 128 * 2 + 4 * 94_208}
# DRAM bytes per launch measured by ncu at nseq = 2048 (final build of round 2: profiles/r02_ncu_full_v8_attention_summary.json,
# profiles/r02_ncu_full_v8_token_summary.json)
NCU_DRAM_BYTES_PER_LAUNCH = {"attention": 772_288_768 + 229_882_624, "token_mid": 889_779_456 + 1_202_538_000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="series per GPU per step")
    ap.add_argument("--length", type=int, default=96)
    ap.add_argument("--rf-steps", type=int, default=100)
    ap.add_argument("--cfg", type=float, default=7.0)
    ap.add_argument("--backbone", default="flowmatching", choices=["flowmatching", "ddpm"])
    ap.add_argument("--chunk", type=int, default=0, help="samples per launch wave (0 = whole batch)")
    ap.add_argument("--cpu-sample", type=int, default=24, help="series in the bounded CPU-baseline sample (~10-15 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference-PyTorch-eager-on-this-GPU baseline leg")
    ap.add_argument("--fused", action="store_true", help="run the guided steps through the fused per-step kernel (csrc/dit_fused.cuh; off by default)")
    ap.add_argument("--strong", action="store_true",
                    help="also time the GLOBAL batch (batch x gpus) on rank 0 alone and add a strong_scaling object (BASELINE config 3)")
    ap.add_argument("--train-batch", type=int, default=256, help="latents per GPU per optimizer step of the training leg (0 = skip)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1399.4))), "hbm": float(d.get("hbm_gbs", 6546.2)),
                "src": "measured (MEASURED_PEAKS.json, bf16/fp16 dense sustained)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _reference_objects(device="cpu"):
    """The unmodified reference modules staged in baseline/_ref (oracle/ref_install.py), or None when they are not there."""
    from oracle import ref_install as R
    from t2ms_b200 import synth
    if not R.available():
        return None
    ref = R.load()
    dit, vae = R.build_reference_models(ref, synth.make_dit_state(0), synth.make_vae_state(1), device)
    return R, ref, dit, vae


def cpu_reference_run(a, n_series: int, repeats: int = 1):
    """The reference loop (infer.py:75-95) on the host cores: the UNMODIFIED reference modules from baseline/_ref
    (kind "reference"), or the oracle port when they are not staged (kind "port").
    Returns (series/s, seconds, threads, kind)."""
    from oracle import t2s_oracle as O
    from t2ms_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    emb, noise = synth.make_text_embeddings(n_series), synth.make_noise(n_series)
    objs = _reference_objects("cpu")
    if objs is not None:
        R, ref, rdit, rvae = objs
        run = lambda: R.reference_sample(ref, rdit, rvae, emb, noise, a.rf_steps, a.cfg, a.length, a.backbone)
        kind = "reference"
    else:
        dsd, vsd = synth.make_dit_state(0), synth.make_vae_state(1)
        if a.backbone == "flowmatching":
            run = lambda: O.rf_sample(dsd, vsd, noise, emb, a.rf_steps, a.cfg, a.length)
        else:
            sn = synth.make_step_noise(a.rf_steps, n_series)
            run = lambda: O.ddpm_sample(dsd, vsd, noise, emb, a.rf_steps, a.cfg, sn, a.length)
        kind = "port"
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_series / best, best, threads, kind


def gpu_eager_baseline(a, dev):
    """the reference PyTorch-eager on the B200 itself — the same unmodified modules from baseline/_ref moved to
    the GPU, fp32, stock settings (torch's default: no TF32 matmuls), the loop of infer.py:75-95 at the bench workload, device
    resident inputs, CUDA events.  Also with TF32 matmuls allowed, and the kernel launches per guided step (torch.profiler)."""
    objs = _reference_objects(dev)
    if objs is None:
        return {"unavailable": "baseline/_ref is not staged (run oracle/ref_install.py where /root/reference is mounted)"}
    R, ref, rdit, rvae = objs
    from t2ms_b200 import synth
    B = a.batch
    emb = synth.make_text_embeddings(B).to(dev)
    noise = torch.randn(B, 64, 30, device=dev)
    res = {"kind": "reference modules (baseline/_ref) on cuda, torch eager fp32", "batch": B, "denoise_steps": a.rf_steps, "unit": UNIT}

    def timed(steps):
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        R.reference_sample(ref, rdit, rvae, emb, noise, steps, a.cfg, a.length, a.backbone)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1)

    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        for allow in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = allow
            timed(3)                                                    # warm-up: cuBLAS / cuDNN / SDPA heuristics, allocator
            ms3 = timed(3)
            # bounded: the whole loop when it fits ~20 s, else extrapolated from a measured prefix of the steps (same per-step work)
            steps = a.rf_steps if ms3 / 3 * a.rf_steps < 20000 else max(3, int(20000 / (ms3 / 3)))
            ms = timed(steps)
            if steps != a.rf_steps:
                ms = ms * a.rf_steps / steps
            key = "value" if not allow else "value_tf32_allowed"
            res[key] = B / (ms / 1e3)
            res["ms_per_batch" if not allow else "ms_per_batch_tf32_allowed"] = ms
            res["measured_steps" if not allow else "measured_steps_tf32_allowed"] = steps
        res["tf32_allowed"] = False
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            R.reference_sample(ref, rdit, rvae, emb[:8], noise[:8], 2, a.cfg, a.length, a.backbone)
            torch.cuda.synchronize(dev)
        n = sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
                and "memset" not in e.name.lower())
        res["launches_per_step"] = n // 2
    except Exception as exc:                                            # CUPTI may be unavailable on the box
        res["launches_per_step"] = None
        res["launches_note"] = f"torch.profiler unavailable: {type(exc).__name__}"
    return res


def workload_name(a):
    kind = "rectified-flow" if a.backbone == "flowmatching" else "DDPM"
    if a.backbone == "flowmatching" and (a.length, a.batch, a.rf_steps) == (96, 1024, 100):
        tag = "BASELINE config 2"
    elif a.backbone == "ddpm" and (a.length, a.rf_steps) == (48, 1000):
        tag = "BASELINE config 3: batch 512 sharded over the GPUs" if a.batch * max(a.gpus, 1) == 512 else "BASELINE config 3 shape"
    else:
        tag = "BASELINE config 5 sweep point"
    return (f"T2S-DiT {kind} sampling length {a.length} with CFG {a.cfg:g}, batch {a.batch}/GPU, {a.rf_steps} steps, "
            f"LA-VAE decode ({tag})")


def run_reference(a, rank, world):
    if rank != 0:
        return
    n = a.cpu_sample
    times = []
    for i in range(a.warmup + a.steps):
        v, dt, threads, kind = cpu_reference_run(a, n)
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": f"{n} series per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} series x {a.rf_steps} guided steps + decode per step, " + BASELINE_WHAT[kind]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def kernel_breakdown(smp, dit, vae, a, dev):
    """Per-kernel device time at the bench workload size, CUDA events on the launching stream."""
    import ctypes as C
    from t2ms_b200 import _lib
    from t2ms_b200.denoiser import _aligned
    lib = _lib.load()
    nseq = 2 * (a.chunk or a.batch)
    pk = dit.packed()
    ws = dit.workspace(nseq, dev)
    wp = _aligned(ws)
    st = torch.cuda.current_stream(dev).cuda_stream
    x = torch.randn(nseq // 2, 64, 30, device=dev)
    emb = torch.randn(nseq // 2, 128, device=dev)
    t100 = torch.full((1,), 37.0, device=dev)
    out = torch.empty(nseq, 64, 30, device=dev)
    series = torch.empty(nseq // 2, a.length, device=dev)
    calls = {
        "cond": lambda: lib.t2s_dit_cond(pk.ref, t100.data_ptr(), 0, emb.data_ptr(), 1, nseq, wp, st),
        "embed_qkv": lambda: lib.t2s_dit_embed_qkv(pk.ref, x.data_ptr(), 1, nseq, wp, st),
        "attention": lambda: lib.t2s_dit_attention(nseq, wp, st),
        "token_mid": lambda: lib.t2s_dit_block_post(pk.ref, 1, nseq, wp, st),
        "token_final": lambda: lib.t2s_dit_final(pk.ref, out.data_ptr(), nseq, wp, st),
        "vae_decode": lambda: vae.decoder.decode_into(x, a.length, series, None),
    }
    res = {}
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        res[name] = e0.elapsed_time(e1) / reps
    return res, nseq


def train_leg(a, dev, rank, world, dist):
    """BASELINE config 4: DiT + frozen LA-VAE training step on mixed lengths 24/48/96 (train.py:52-90), data-parallel:
    one optimizer step per length-grouped sub-batch = encoder, create_flow, CFG-drop coin, forward, MSE, backward,
    NCCL all-reduce of the flat gradient bucket, AdamW.  Returns optimizer steps/s (max over ranks timing)."""
    from t2ms_b200 import Transformer, synth, vqvae
    from t2ms_b200.compat import VAE_ARGS
    from t2ms_b200.training import DitTrainer
    dit = Transformer()
    dit.load_state_dict(synth.make_dit_state(0))
    dit = dit.to(dev).train()
    vae = vqvae(VAE_ARGS)
    vae.load_state_dict(synth.make_vae_state(1))
    vae = vae.to(dev).eval()
    tr = DitTrainer(dit)
    B = a.train_batch
    g = torch.Generator().manual_seed(77 + rank)
    batches = [(torch.rand(B, L, generator=g).to(dev), torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=-1).to(dev))
               for L in (24, 48, 96)]

    def one_pass():
        for series, emb in batches:
            tr.train_batch(series, emb, backbone="flowmatching", total_step=100, encoder=vae.encoder)

    for _ in range(2):
        one_pass()
    reps = 3
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        one_pass()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    nsteps = reps * len(batches)
    sps = nsteps / (ms.item() / 1e3)
    return {"metric": "t2s_dit_train_steps_per_sec", "value": sps, "unit": "optimizer steps/s", "samples_per_sec": sps * B * world,
            "ms_per_step": ms.item() / nsteps, "batch_per_gpu": B, "global_batch": B * world, "lengths": [24, 48, 96],
            "dtype": "tf32 operands (Linears) and f16 operands (fused attention) / f32 accumulate, fp32 master weights and AdamW state",
            "workload": "BASELINE config 4: frozen LA-VAE encode + rectified-flow training step, mixed lengths 24/48/96, "
                        f"data-parallel x{world} (one flat-bucket NCCL all-reduce per step)",
            "tflops_algorithmic": sps * B * world * 3 * FLOP_FWD / 1e12, "final_loss": float(tr.loss_sum.item() / (B * 1920 * world))}


def main():
    a = parse()
    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    import torch.distributed as dist
    from t2ms_b200 import T2SSampler, Transformer, synth, vqvae
    from t2ms_b200.compat import VAE_ARGS
    from t2ms_b200.sampler import gather_series

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    dit = Transformer()
    dit.load_state_dict(synth.make_dit_state(0))
    dit = dit.to(dev).eval()
    vae = vqvae(VAE_ARGS)
    vae.load_state_dict(synth.make_vae_state(1))
    vae = vae.to(dev).eval()
    smp = T2SSampler(dit, vae)
    if a.fused:
        from t2ms_b200 import _lib
        _lib.load().t2s_set_fused(1, 0)
    B = a.batch
    gen = torch.Generator().manual_seed(1234 + rank)
    emb_host = torch.nn.functional.normalize(torch.randn(B, 128, generator=gen), dim=-1).pin_memory()
    out_host = torch.empty(B, a.length, dtype=torch.float32).pin_memory()
    emb_dev = emb_host.to(dev)
    noise = torch.randn(B, 64, 30, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    kw = dict(steps=a.rf_steps, cfg_scale=a.cfg, backbone=a.backbone, chunk=a.chunk or None)

    def step_device():
        flush.zero_()                                                   # L2 flush between timed iterations
        s = smp.sample(emb_dev, a.length, noise=noise, **kw)
        return gather_series(s, B * world) if world > 1 else s

    def step_e2e():
        flush.zero_()
        s = smp.sample_host(emb_host, a.length, out_host=out_host, **kw)
        return s

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(a.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_device, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)

    value = world * B * a.steps / (ms / 1e3)
    e2e = world * B * a.steps / (ms_e2e / 1e3)
    strong = None
    if a.strong and world > 1:
        # strong scaling (BASELINE config 3: a fixed global batch sharded over the GPUs): the same global batch on ONE GPU, timed
        # on rank 0 while the other ranks wait; efficiency = rate_N / (N x rate_1) at equal global batch
        if rank == 0:
            Bg = B * world
            emb_g = torch.nn.functional.normalize(torch.randn(Bg, 128, generator=gen), dim=-1).to(dev)
            noise_g = torch.randn(Bg, 64, 30, device=dev)
            smp.sample(emb_g, a.length, noise=noise_g, **kw)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            smp.sample(emb_g, a.length, noise=noise_g, **kw)
            e1.record()
            torch.cuda.synchronize(dev)
            one = Bg / (e0.elapsed_time(e1) / 1e3)
            strong = {"global_batch": Bg, "batch_per_gpu": B, "n_gpus": world, "value": value, "unit": UNIT, "single_gpu_value": one,
                      "efficiency": value / (world * one),
                      "note": "same global batch on one GPU of the same box (rank 0, after the timed region); the loss is wave quantisation "
                              "of small per-GPU batches on 148 SMs, not communication (no collective inside the loop)"}
        dist.barrier()
    del flush
    train = train_leg(a, dev, rank, world, dist) if a.train_batch > 0 else None
    line = None
    if rank == 0:
        pk = peaks()
        kb, nseq = kernel_breakdown(smp, dit, vae, a, dev)
        n_chunks = 1 if not a.chunk else -(-B // a.chunk)
        step_kernel_ms = a.rf_steps * (kb["cond"] + kb["embed_qkv"] + 4 * kb["attention"] + 3 * kb["token_mid"] + kb["token_final"]) * n_chunks
        shares = {k: round(a.rf_steps * n_chunks * v * {"attention": 4, "token_mid": 3}.get(k, 1) / step_kernel_ms, 4)
                  for k, v in kb.items() if k != "vae_decode"}
        dom = max(("attention", "token_mid"), key=lambda k: shares[k])
        flop = (FLOP_ATTN_LAYER if dom == "attention" else FLOP_TOKEN_MID) * nseq
        achieved = flop / (kb[dom] * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (10-bit mantissa = tf32 operand precision); residual, LayerNorm, softmax, update in f32",
            "data": "synthetic (random-init weights seed 0/1, unit-norm 128-d text embeddings, Gaussian initial latents)",
            "config": {"workload": workload_name(a), "batch_per_gpu": B, "length": a.length, "denoise_steps": a.rf_steps,
                       "cfg_scale": a.cfg, "backbone": a.backbone, "parallelism": f"batch-shard x{world}, final all_gather",
                       "l2": "256 MiB flush write between timed iterations; per-step working set (1.5 GB scratch) exceeds the 126 MB L2"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": emb_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": a.steps * (a.rf_steps * (2 if a.fused else 10) * n_chunks + 1),
            "clocks": clocks,
            "tflops_algorithmic": world * B * a.steps * (2 * a.rf_steps * FLOP_FWD + FLOP_DECODE_96) / (ms / 1e3) / 1e12,
            "roofline": {"bound": "tensor", "kernel": "attn_kernel" if dom == "attention" else "token_kernel<MID>",
                         "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(dom) if nseq == 2048 else None,
                         "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum at this workload size "
                                           "(profiles/r02_ncu_full_v8_attention_summary.json, r02_ncu_full_v8_token_summary.json)",
                         "algorithmic_bytes": ALGO_BYTES_PER_SEQ[dom] * nseq,
                         "co_bound": "MUFU ex2: 0.9216 M exp per sequence-block at 16/clk/SM = 0.41 ms per 2048-sequence launch "
                                     "(ncu: XU pipe 79 % busy, tensor pipe 19 %); this kernel is bound by the MUFU, not the tensor pipe"
                                     if dom == "attention" else "epilogue issue",
                         "peak_source": pk["src"], "flop_per_launch": flop, "launch_ms": kb[dom]},
            # what actually bounds the dominant kernel: 480 x 480 exponentials per (sequence, head) on the 16 MUFU lanes per SM
            "mufu_floor": ({"exp_per_launch": nseq * 4 * 480 * 480, "floor_ms": nseq * 4 * 480 * 480 / (16 * 148 * ((clocks or {}).get("sm_mhz") or 1837.0) * 1e3),
                            "launch_ms": kb["attention"],
                            "frac": nseq * 4 * 480 * 480 / (16 * 148 * ((clocks or {}).get("sm_mhz") or 1837.0) * 1e3) / kb["attention"],
                            "note": "attention kernel: ex2 count / (16 per clock per SM x 148 SMs x the SM clock sampled during the timed region)"}),
            "kernel_ms": {k: round(v, 4) for k, v in kb.items()},
            "kernel_share_of_step": shares,
        }
        line["train"] = train
        if strong is not None:
            line["strong_scaling"] = strong
        # the memory-bound tail of the path: LA-VAE decode (one launch per batch), reported against the measured HBM peak
        dec_ms = kb["vae_decode"]
        dec_bytes = (nseq // 2) * (64 * 30 * 4 + a.length * 4) + 319_937 * 4
        line["vae_decode"] = {"ms": dec_ms, "series": nseq // 2, "algorithmic_bytes": dec_bytes, "achieved_gbs": dec_bytes / (dec_ms * 1e-3) / 1e9,
                              "hbm_peak_gbs": pk["hbm"], "frac_of_hbm_peak": dec_bytes / (dec_ms * 1e-3) / 1e9 / pk["hbm"],
                              "tflops_fp32": (nseq // 2) * 640_000 * (a.length // 4) / (dec_ms * 1e-3) / 1e12,
                              "bound": "latency / L2 weight reads (1.28 MB of fp32 weights per group of series), 0.1 % of the workload's time"}
        if not a.no_cpu_baseline and world == 1:       # reported at N=1 only (host cores are shared by the ranks otherwise)
            v, dt, threads, kind = cpu_reference_run(a, a.cpu_sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"{a.cpu_sample} series x {a.rf_steps} guided steps + decode in {dt:.1f} s, " + BASELINE_WHAT[kind]}
        if not a.no_gpu_eager and world == 1:
            line["gpu_eager_baseline"] = gpu_eager_baseline(a, dev)
            if line["gpu_eager_baseline"].get("value"):
                line["gpu_eager_baseline"]["ours_over_eager"] = value / line["gpu_eager_baseline"]["value"]
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
