#!/usr/bin/env python
"""BASELINE config 5: throughput sweep batch x length x denoising steps (x GPUs under torchrun) of the guided
rectified-flow sampler through the host-buffer entry (`T2SSampler.sample_host`: pinned text embeddings in, series out),
next to the reference loop (infer.py:75-95; the unmodified modules from baseline/_ref) on the host cores for the sizes it finishes in seconds.

    python tools/sweep.py [--out gpurun_out/sweep.json] [--batches 1,8,...] [--cpu-budget-s 40]
    python -m torch.distributed.run --nproc-per-node N ... tools/sweep.py     (weak scaling: every rank runs `batch`)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--batches", default="1,8,64,512,1024,4096,8192")
    ap.add_argument("--lengths", default="24,48,96")
    ap.add_argument("--steps", default="10,50,100")
    ap.add_argument("--backbone", default="flowmatching")
    ap.add_argument("--cfg", type=float, default=7.0)
    ap.add_argument("--cpu-budget-s", type=float, default=40.0, help="total host time spent on the CPU reference legs (0 = skip)")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch.distributed as dist
    from t2ms_b200 import T2SSampler, Transformer, synth, vqvae
    from t2ms_b200.compat import VAE_ARGS
    from t2ms_b200.sampler import gather_series
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    dit = Transformer()
    dit.load_state_dict(synth.make_dit_state(0))
    dit = dit.to(dev).eval()
    vae = vqvae(VAE_ARGS)
    vae.load_state_dict(synth.make_vae_state(1))
    vae = vae.to(dev).eval()
    smp = T2SSampler(dit, vae)
    rows = []
    for B in [int(x) for x in a.batches.split(",")]:
        g = torch.Generator().manual_seed(99 + rank)
        emb_host = torch.nn.functional.normalize(torch.randn(B, 128, generator=g), dim=-1).pin_memory()
        for L in [int(x) for x in a.lengths.split(",")]:
            out_host = torch.empty(B, L, dtype=torch.float32).pin_memory()
            for steps in [int(x) for x in a.steps.split(",")]:
                def run():
                    s = smp.sample_host(emb_host, L, out_host=out_host, steps=steps, cfg_scale=a.cfg, backbone=a.backbone)
                    if world > 1:
                        gather_series(s.to(dev), B * world)
                run()                                                    # warm-up (allocations, tables)
                reps = 3 if B * steps <= 64 * 100 else 1
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize(dev)
                ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                rows.append({"batch_per_gpu": B, "n_gpus": world, "length": L, "steps": steps, "ms": round(ms.item(), 3),
                             "series_per_s": round(B * world / (ms.item() / 1e3), 2)})
    cpu = []
    if rank == 0 and a.cpu_budget_s > 0:
        # the CPU reference legs are bench.py's own cpu_baseline leg (the oracle port of infer.py:75-95 on all host cores)
        import bench
        from argparse import Namespace
        spent = 0.0
        for B, L, steps in ((1, 24, 10), (8, 24, 10), (8, 96, 10), (8, 24, 50), (8, 96, 100), (64, 96, 10)):
            if spent > a.cpu_budget_s:
                break
            v, dt, threads, kind = bench.cpu_reference_run(Namespace(backbone=a.backbone, rf_steps=steps, cfg=a.cfg, length=L), B)
            spent += dt
            cpu.append({"batch": B, "length": L, "steps": steps, "s": round(dt, 3), "series_per_s": round(v, 3), "cores": threads,
                        "kind": kind + ": " + bench.BASELINE_WHAT[kind]})
    if rank == 0:
        out = {"workload": "BASELINE config 5: guided RF sampling sweep, host buffers in/out (e2e), CFG %g" % a.cfg, "n_gpus": world,
               "gpu": rows, "cpu_reference": cpu}
        txt = json.dumps(out, indent=1)
        if a.out:
            with open(a.out, "w") as f:
                f.write(txt)
        print(txt)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
