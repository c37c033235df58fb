import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from gpu_util import DEV, make_dit, make_vae
from t2ms_b200 import T2SSampler, synth
(dit, _), (vae, _) = make_dit(3), make_vae(4)
smp = T2SSampler(dit, vae)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
emb, x0 = synth.make_text_embeddings(B, seed=5).to(DEV), synth.make_noise(B, seed=6).to(DEV)
for _ in range(3):
    smp.sample(emb, 96, steps=3, noise=x0)
torch.cuda.synchronize()
