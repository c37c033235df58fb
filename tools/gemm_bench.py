#!/usr/bin/env python
"""Device time of the training GEMM shapes (t2s_gemm_tf32, C = A B^T) against their HBM roofline, with the persistent
kernel's profiling knobs (mode bit 16: no global stores, bit 32: no loads) to see which role bounds a launch."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200 import _lib

DEV = "cuda:0"
lib = _lib.load()
T = 256 * 480
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6546.2
st = torch.cuda.current_stream().cuda_stream
rows = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, M, N, K, b_mn in (("qkv fwd", T, 384, 128, 0), ("proj fwd", T, 128, 128, 0), ("fc1 fwd", T, 256, 128, 0), ("fc2 fwd", T, 128, 256, 0),
                            ("fc2 dX", T, 256, 128, 1), ("fc1 dX", T, 128, 256, 1), ("qkv dX", T, 128, 384, 1)):
    A = torch.randn(M, K, device=DEV)
    B = torch.randn(K, N, device=DEV) if b_mn else torch.randn(N, K, device=DEV)
    C = torch.empty(M, N, device=DEV)
    bias = torch.randn(N, device=DEV)
    r = {"gemm": name, "M": M, "N": N, "K": K, "b_mn": b_mn}
    for label, mode in (("ms", 0), ("ms_no_store", 16), ("ms_no_load", 32), ("ms_neither", 48)):
        def run():
            _lib.check(lib.t2s_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), bias.data_ptr(), M, N, K, K, N if b_mn else K, N, 0, b_mn, mode, 1.0, 1, st))
        for _ in range(2):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        r[label] = round(sorted(ts)[2], 4)
    byts = 4 * (M * K + M * N + N * K)
    r["roofline_ms"] = round(byts / PEAK / 1e6, 4)
    r["achieved_gbs"] = round(byts / r["ms"] / 1e6, 1)
    r["frac"] = round(r["roofline_ms"] / r["ms"], 3)
    rows.append(r)
print(json.dumps(rows, indent=1))
