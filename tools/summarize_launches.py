#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares of the LAST
`--steps`-th part of the run (the earlier parts are warm-up).  Usage: summarize_launches.py file.csv [parts] > summary.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            v = float(r[idx["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[idx["Metric Unit"]]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        recs.append((r[idx["Kernel Name"]], v, r[idx["Grid Size"]], r[idx["Block Size"]]))
    n = len(recs)
    last = recs[n - n // parts:]
    tot = sum(r[1] for r in last)
    print(f"source: {path}\n\nlaunches in the measured part: {len(last)}; summed kernel time {tot / 1e3:.2f} ms "
          f"(ncu serialises launches and runs them cold-cache: compare SHARES, not absolutes)\n")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v, _, _ in last:
        k = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "")[:60]
        agg[k][0] += 1
        agg[k][1] += v
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        print(f"| `{k}` | {c} | {v:.1f} | {100 * v / tot:.1f} % |")
    print("\nTop individual launches:\n\n| us | kernel | grid | block |\n|---:|---|---|---|")
    for name, v, g, b in sorted(last, key=lambda r: -r[1])[:15]:
        print(f"| {v:.1f} | `{name[:70]}` | {g} | {b} |")


if __name__ == "__main__":
    main()
