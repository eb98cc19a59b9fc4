#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share,
and (optionally) the ordered launch sequence of one decoder pass.   python tools/launch_summary.py launches.csv [--seq]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, v, row["Grid Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    for name, v, _ in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {len(rows)} launches, sum of kernel durations {tot:.2f} ms\n")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:70]} | {a[0]} | {a[1]:.3f} | {1e3 * a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% |")
    if "--seq" in sys.argv:
        idx = [i for i, r in enumerate(rows) if "dense_grid" in r[0]]
        if len(idx) >= 3:
            print("\n# one decoder pass (launch order)")
            for name, v, grid in rows[idx[1]:idx[2]]:
                print(f"{name[:50]:50s} {grid:16s} {v * 1e3:9.1f} us")
    if "--enc" in sys.argv:
        idx = [i for i, r in enumerate(rows) if "dense_grid" in r[0]]
        end = idx[0] if idx else len(rows)
        print("\n# encoder launches (before the first decoder pass)")
        for name, v, grid in rows[:end]:
            print(f"{name[:50]:50s} {grid:16s} {v * 1e3:9.1f} us")


if __name__ == "__main__":
    main()
