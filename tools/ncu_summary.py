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
    name = d[idx["Kernel Name"]].split("(")[0].replace("zs::", "").replace("void ", "").split("<")[0].strip()
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
    # launches of one decoder pass (round-2 kernels): point_proj, 2 x (LN+qkv+attention, proj+residual, MLP), occupancy MLP.
    # A capture that misses a launch of a repeated kernel reuses the captured one (same shapes); a missing point_proj is
    # counted with its algorithmic bytes (12 B read + 1024 B written per point).
    label = {"chain_qkvattn2_kernel": "chain_qkvattn", "chain_qkvattn_kernel": "chain_qkvattn", "chain_lin_kernel": "chain_lin[proj]",
             "chain_mlp2_kernel": "chain_mlp", "chain_mlp_kernel": "chain_mlp", "chain_pmlp_kernel": "chain_pmlp",
             "chain_occ2_kernel": "chain_occ", "chain_occ_kernel": "chain_occ", "point_proj_kernel": "point_proj",
             "chain_attn_kernel": "attn_fused"}
    per_pass = {"chain_qkvattn": 2, "chain_lin[proj]": 2, "chain_mlp": 2, "chain_pmlp": 2, "chain_occ": 1, "point_proj": 1, "attn_fused": 2}
    per = {}
    for k, v in traffic.items():
        if k in label:
            per[label[k]] = sum(v) / len(v) / pts          # mean over the captured launches of that kernel
    assumed = []
    if "point_proj" not in per and "--no-point-proj" not in sys.argv:
        per["point_proj"] = 1036.0
        assumed.append("point_proj (not captured: algorithmic 1036 B/pt)")
    chain = sum(v * per_pass[k] for k, v in per.items())
    json.dump({"source": f"{rep} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of a {pts}-point pass; tools/ncu_summary.py)",
               "points_per_profiled_launch": pts, "bytes_per_point": per, "launches_per_pass": {k: per_pass[k] for k in per},
               "assumed": assumed, "chain": chain}, open(sys.argv[3], "w"), indent=1)
