#!/usr/bin/env python
"""Hot instructions of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > file.csv`:
prints the SASS lines with the most warp-stall samples and their two dominant stall reasons."""
import csv
import sys


def main(path, top_n=40):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Address" in r and "# Samples" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows:
        if len(r) != len(hdr) or r is hdr:
            continue
        try:
            int(r[ix["# Samples"]])
        except ValueError:
            continue
        data.append(r)
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"{len(data)} instructions, {tot} samples")
    agg = {}
    for r in data:
        for s in stalls:
            agg[s] = agg.get(s, 0) + int(r[ix[s]] or 0)
    print("by reason:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for i, r in enumerate(data):
        r.append(i)
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
        n = int(r[ix["# Samples"]])
        st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"#{r[-1]:5d} {n:6d} {100 * n / tot:5.1f}%  {r[ix['Source']][:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
