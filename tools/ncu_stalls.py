#!/usr/bin/env python3
"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "# Samples" in r)
isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows:
    if len(r) == len(hdr) and r[isamp].isdigit():
        data.append(r + [len(data)])
    elif data and r and r[0] == "Address":
        break                      # second view (PTX / CUDA-C) starts: keep the SASS view only
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "SASS lines", len(data))
for r in sorted(data, key=lambda r: -int(r[isamp]))[:n]:
    st = sorted(((hdr[i], int(r[i])) for i in stall_cols if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{r[-1]:5d} {int(r[isamp]):6d} {r[iex]:>8s}  {r[isrc].strip()[:78]:78s} {st}")
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
