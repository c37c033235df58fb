"""Build a variant of the library with extra nvcc flags for A/B runs: python tools/build_variant.py NAME -DFLAG[=V] ...
-> t2ms_b200/lib/variants/libt2s_b200_NAME.so (select with T2S_B200_LIB=<path>)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t2ms_b200.build import CSRC, NVCC_FLAGS, SOURCES  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "t2ms_b200", "lib", "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libt2s_b200_{name}.so")
cmd = ["nvcc", *NVCC_FLAGS, *flags, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode != 0:
    sys.exit(r.stdout + r.stderr)
print(out)
