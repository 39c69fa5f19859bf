#!/usr/bin/env python
"""Stall-sample hot spots of one kernel from `ncu --page source --csv` (SASS view): instructions in address order with
their share of samples, executed count and dominant stall reason; consecutive cold instructions are folded.

  ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > k.csv ; python profiles/hotspots.py k.csv [min_pct]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["Address"]].startswith("0x")]
seen, uniq = set(), []
for r in body:
    if r[ix["Address"]] in seen:
        continue
    seen.add(r[ix["Address"]])
    uniq.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in uniq)
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
print("total samples", tot, "instructions", len(uniq))
cold_s = cold_n = 0
for r in uniq:
    s = int(r[ix["# Samples"]])
    pct = 100.0 * s / max(tot, 1)
    if pct < minpct:
        cold_s += s
        cold_n += 1
        continue
    if cold_n:
        print("   ... %d instructions, %.2f%% of samples" % (cold_n, 100.0 * cold_s / tot))
        cold_s = cold_n = 0
    top = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print("%s %5.2f%% exec %10s  %-28s %s" % (r[ix["Address"]][-5:], pct, r[ix["Instructions Executed"]],
                                            ",".join("%s:%d" % (h[6:], v) for v, h in top if v), r[ix["Source"]][:70]))
if cold_n:
    print("   ... %d instructions, %.2f%% of samples" % (cold_n, 100.0 * cold_s / tot))
