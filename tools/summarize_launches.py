#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1; agg[n][1] += ns
    print(f"launches {len(rows)}  total {tot/1e6:.3f} ms (serialised, cold-cache: compare shares)")
    print(f"{'share':>6} {'ms':>9} {'count':>6} {'avg us':>9}  kernel")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*ns/tot:6.2f} {ns/1e6:9.3f} {c:6d} {ns/c/1e3:9.1f}  {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
