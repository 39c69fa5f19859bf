#!/usr/bin/env python
"""Dynamic instruction mix and stall-sample split of one kernel from `ncu --page source --csv` output.

  ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > k.csv ; python profiles/srcmix.py k.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["Address"]].startswith("0x")]
seen, uniq = set(), []
for r in body:           # the page lists the kernel twice (SASS, then source-correlated): keep first occurrence
    if r[ix["Address"]] in seen:
        continue
    seen.add(r[ix["Address"]])
    uniq.append(r)
mix, smp = collections.Counter(), collections.Counter()
tot = stot = 0
for r in uniq:
    toks = r[ix["Source"]].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0].rstrip(";")
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    mix[op] += n
    smp[op] += s
    tot += n
    stot += s
print("warp instructions executed: %.3f G, samples %d" % (tot / 1e9, stot))
print("%-10s %8s %8s" % ("opcode", "% instr", "% samples"))
for k, v in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print("%-10s %7.2f%% %7.2f%%" % (k, 100 * v / tot, 100 * smp[k] / max(stot, 1)))
