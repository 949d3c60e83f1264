#!/usr/bin/env python
"""Top stall-sampled SASS instructions of an `ncu --page source --csv` export, with +-ctx lines of context."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
body = rows[hi + 1:]
samp = [float(r[idx["# Samples"]] or 0) for r in body]
tot = sum(samp)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
order = sorted(range(len(body)), key=lambda i: -samp[i])[:top]
for i in order:
    r = body[i]
    st = sorted(((float(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{samp[i]:7.0f} {100*samp[i]/tot:5.1f}% line {i:5d} exec {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:90]:90s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
    if ctx:
        for j in range(max(0, i - ctx), min(len(body), i + ctx + 1)):
            print(f"        {j:5d} {samp[j]:6.0f} {body[j][idx['Source']].strip()[:100]}")
