"""A few guided steps through the fused per-step kernel (for ncu captures): python tools/fused_once.py [batch] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2ms_b200 import T2SSampler, Transformer, _lib, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fused = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = "cuda:0"
lib = _lib.load()
lib.t2s_set_fused(1 if fused else -1, 0)
dit = Transformer(); dit.load_state_dict(synth.make_dit_state(0)); dit = dit.to(dev).eval()
smp = T2SSampler(dit)
emb, x0 = synth.make_text_embeddings(B, seed=7).to(dev), torch.randn(B, 64, 30, device=dev)
lat = smp.sample_latent(emb, steps=steps, noise=x0)
torch.cuda.synchronize()
print("ok", float(lat.abs().max()))
