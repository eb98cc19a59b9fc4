#!/usr/bin/env python
"""Aggregate the per-instruction warp-stall samples of one kernel from an ncu report (SASS view).
    python tools/ncu_stalls.py report.ncu-rep <kernel-regex> [launch-skip]
Prints the stall-reason totals and the 25 hottest SASS instructions."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
hdr, data = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
recs = []
for d in data:
    if len(d) < len(hdr):
        continue
    n = int(d[ix["# Samples"]] or 0)
    for h in stall_cols:
        tot[h] += int(d[ix[h]] or 0)
    recs.append((n, d[ix["Source"]].strip(), {h: int(d[ix[h]] or 0) for h in stall_cols}))
total = sum(r[0] for r in recs)
print(lines[0])
print("total samples", total)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d}  {100.0 * v / max(total, 1):5.1f}%")
print("hottest instructions:")
for n, src, st in sorted(recs, key=lambda r: -r[0])[:25]:
    top = max(st.items(), key=lambda kv: kv[1])
    print(f"  {n:7d} {100.0 * n / max(total, 1):5.1f}%  {src[:70]:70s} {top[0]}")
