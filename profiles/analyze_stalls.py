"""Warp-state (stall) summary of one kernel from an `ncu --set full --import-source on` report.

usage: ncu -i report.ncu-rep --page source --csv --print-source sass > sass.csv; python profiles/analyze_stalls.py sass.csv
Prints the share of every stall reason over all samples and the hottest instructions with their top reasons."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, samples, top = collections.Counter(), 0, []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    samples += n
    st = {s: int(r[idx[s]] or 0) for s in stalls if (r[idx[s]] or "0") != "0"}
    tot.update(st)
    top.append((n, r[idx["Source"]][:72], st))
print("samples", samples)
for s, v in tot.most_common(12):
    print(f"{s:26s} {v:8d} {100 * v / samples:5.1f}%")
for n, src, st in sorted(top, key=lambda x: -x[0])[:20]:
    print(f"{n:7d} {100 * n / samples:5.1f}%  {src:72s} {sorted(st.items(), key=lambda kv: -kv[1])[:2]}")
