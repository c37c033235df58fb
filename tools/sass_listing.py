"""Committed SASS evidence (north_star: "a committed SASS listing"): the disassembly of the hot kernels of the built library,
instruction text only (addresses and encodings stripped), plus the mnemonic histogram of each.
    python tools/sass_listing.py profiles/r02_sass"""
import collections
import re
import subprocess
import sys

LIB = "t2ms_b200/lib/libt2s_b200.so"
KERNELS = {
    "attn_kernel_H30": "_ZN3t2s11attn_kernelILi30ELb0EEEvPK6__halfPS1_Pxi",
    "token_kernel_MID_H30_NE2": "_ZN3t2s12token_kernelILi1ELi30ELi2EEEvNS_7TokArgsE",
    "fused_step_kernel": "_ZN3t2s17fused_step_kernelENS_9FusedArgsE",
}
out_prefix = sys.argv[1]
for name, sym in KERNELS.items():
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    ins = [m.group(1).strip() for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);", txt, re.M)]
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
    with open(f"{out_prefix}_{name}.txt", "w") as f:
        f.write(f"# cuobjdump -sass -fun {sym} {LIB}\n# {len(ins)} instructions; mnemonic histogram:\n")
        f.write("# " + ", ".join(f"{k} {v}" for k, v in hist.most_common(40)) + "\n")
        f.write("# tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, cp.async.bulk = UBLKCP, mbarrier = SYNCS, ex2 = MUFU.EX2\n")
        f.write("\n".join(ins) + "\n")
    print(name, len(ins), {k: hist[k] for k in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "MUFU", "SYNCS", "HMMA") if k in hist})
