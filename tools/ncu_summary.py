#!/usr/bin/env python
"""Summarise ncu captures (run in the build container, no GPU needed):
    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep  > profiles/rNN_ncu_full_summary.json
    python tools/ncu_summary.py list  gpurun_out/launches.csv  > profiles/rNN_ncu_launch_table.txt
    python tools/ncu_summary.py sass  t2ms_b200/lib/libt2s_b200.so > profiles/rNN_sass_mnemonics.txt
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    print(json.dumps(res, indent=1))


def launch_list(path):
    hdr, data = None, []
    for r in csv.reader(open(path)):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
        elif hdr and r and r[0].isdigit():
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
        a = agg.setdefault(re.sub(r"\(.*", "", d["Kernel Name"])[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(data)} launches, {tot:.1f} us in total (per-launch times under ncu are cold-cache and serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[1]:10.1f} us {100 * a[1] / tot:5.1f} %  x{a[0]:4d}  {a[1] / a[0]:9.1f} us/launch  {k}")


def sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    pat = re.compile(r"\b(UTC[A-Za-z0-9.]*MMA[A-Za-z0-9._]*|UTCBAR[A-Za-z0-9._]*|UTCATOMSWS[A-Za-z0-9._]*|LDTM[A-Za-z0-9._]*|STTM[A-Za-z0-9._]*|UBLKCP[A-Za-z0-9._]*|UBLKPF[A-Za-z0-9._]*|"
                     r"SYNCS[A-Za-z0-9._]*|LDGSTS[A-Za-z0-9._]*|MUFU[A-Za-z0-9._]*|UTMALDG[A-Za-z0-9._]*|HMMA[A-Za-z0-9._]*)")
    fn, counts, n = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            n[fn] = 0
            continue
        if fn and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            n[fn] += 1
            for t in pat.findall(line):
                counts[fn][t] += 1
    print("SASS evidence (cuobjdump -sass): counts of Blackwell-specific mnemonics per kernel")
    print("UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier, LDGSTS = cp.async\n")
    for fn, c in counts.items():
        print(f"{fn}  ({n[fn]} instructions)")
        for k in sorted(c):
            print(f"    {k:32s} {c[k]}")


if __name__ == "__main__":
    {"full": full, "list": launch_list, "sass": sass}[sys.argv[1]](sys.argv[2])
