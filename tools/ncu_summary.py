#!/usr/bin/env python
"""Markdown table of the key ncu metrics of every launch in a report + per-point DRAM traffic JSON.
    python tools/ncu_summary.py report.ncu-rep points_per_launch [traffic.json]"""
import csv
import json
import subprocess
import sys

rep, pts = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) %"), ("launch__registers_per_thread", "regs"),
        ("sm__cycles_elapsed.avg.per_second", "SM GHz")]


def num(d, c):
    v, u = d[idx[c]].replace(",", ""), units[idx[c]]
    f = float(v)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u)
    return f * scale if scale else f


print("| # | kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|---|" + "---:|" * len(cols))
traffic = {}
for n, d in enumerate(data):
    name = d[idx["Kernel Name"]].split("(")[0].replace("zs::", "")
    vals = []
    for c, _ in cols:
        v, u = d[idx[c]], units[idx[c]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.3g}" if f < 1000 else f"{f:.0f}"
        except ValueError:
            pass
        vals.append(f"{v} {u}".replace(" %", "%").replace("register/thread", "").strip())
    print(f"| {n} | {name} | " + " | ".join(vals) + " |")
    traffic.setdefault(name, []).append(num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"))
if len(sys.argv) > 3:
    lin = traffic.get("chain_lin_kernel", [])
    per = {"chain_lin[qkv]": lin[0] / pts if lin else None, "chain_lin[proj]": lin[1] / pts if len(lin) > 1 else None,
           "attn_fused": traffic["chain_attn_kernel"][0] / pts, "chain_mlp": traffic["chain_mlp_kernel"][0] / pts,
           "chain_occ": traffic["chain_occ_kernel"][0] / pts}
    chain = sum(v * (1 if k == "chain_occ" else 2) for k, v in per.items())
    json.dump({"source": f"{rep} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of a {pts}-point pass; tools/ncu_summary.py)",
               "points_per_profiled_launch": pts, "bytes_per_point": per, "chain": chain}, open(sys.argv[3], "w"), indent=1)
