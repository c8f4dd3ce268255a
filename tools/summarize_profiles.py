#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into small tracked summaries under profiles/.
  python tools/summarize_profiles.py <tag> [--launches gpurun_out/launches_X.csv] [--rep name=gpurun_out/prof_X.ncu-rep ...]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])[:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list of `{path}`\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live event timing\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:72s} {n:8d} {t:12.1f} {t / n:10.1f} {100 * t / tot:6.1f}%\n")
    print("wrote", out)


def rep(name, path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none capture `{path}` (one block per captured launch)\n")
        for r in rows[2:]:
            f.write(f"\n[{name}] {r[hdr.index('Kernel Name')][:110]}\n")
            for m in KEEP:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"  {m:95s} {r[i]:>16s} {units[i]}\n")
    print("wrote", out)


if __name__ == "__main__":
    tag = sys.argv[1]
    args = sys.argv[2:]
    i = 0
    while i < len(args):
        if args[i] == "--launches":
            launches(args[i + 1], f"profiles/{tag}_launches.txt")
            i += 2
        elif args[i] == "--rep":
            n, p = args[i + 1].split("=", 1)
            rep(n, p, f"profiles/{tag}_ncu_{n}.txt")
            i += 2
        else:
            raise SystemExit("bad arg " + args[i])
