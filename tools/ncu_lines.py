#!/usr/bin/env python
"""Aggregates the warp-stall samples of an .ncu-rep by CUDA source line (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py report.ncu-rep lib.so kernel_substring [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile, glob
rep, lib, kern = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
iS, iE = h.index("# Samples"), h.index("Instructions Executed")
base = min(int(r[0], 16) for r in data)
samp = {int(r[0], 16) - base: (int(r[iS] or 0), int(r[iE] or 0), r[1]) for r in data}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for f in glob.glob(tmp + "/*.cubin"):
    out = subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout
    if kern not in out:
        continue
    cur = None; active = False
    for l in out.splitlines():
        if l.startswith("\t.section") or "//-----" in l:
            active = kern in l and ".text." in l if ".text." in l else active
        m = re.search(r'//## File ".*?([\w\.]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
        if m and active:
            line_of[int(m.group(1), 16)] = cur
tot = sum(v[0] for v in samp.values())
print("total samples", tot, "instructions", len(samp), "executed", sum(v[1] for v in samp.values()))
agg = {}
for off, (s, e, _) in samp.items():
    k = line_of.get(off)
    a = agg.setdefault(k, [0, 0]); a[0] += s; a[1] += e
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(k, v)
