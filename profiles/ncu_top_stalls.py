#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: total samples by stall reason and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and (r[ix["# Samples"]] or "0").isdigit()]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
allsamp = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", allsamp)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-24s %7d  %5.1f%%" % (s, v, 100.0 * v / max(1, allsamp)))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    top = sorted(stalls, key=lambda s: -int(r[ix[s]] or 0))[:2]
    print("  %6s  %-70s %s" % (r[ix["# Samples"]], r[ix["Source"]][:70], ", ".join("%s=%s" % (t, r[ix[t]]) for t in top)))
