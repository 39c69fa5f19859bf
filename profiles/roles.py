#!/usr/bin/env python
"""Per-role totals of the warp-specialised tensor-core ALS kernel from `ncu --page source --csv` (SASS view): the code of
the roles is laid out in source order (prologue, producers, MMA issue, convert, epilogue), separated here by marker opcodes.
  python profiles/roles.py k.csv
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
src = [r[ix["Source"]] for r in uniq]
first = lambda pat: next(i for i, s in enumerate(src) if pat in s)   # noqa: E731
last = lambda pat: max(i for i, s in enumerate(src) if pat in s)     # noqa: E731
b_prod = first("UBLKCP") - 400 if first("UBLKCP") > 400 else 0
marks = [("prologue", 0), ("producers", None), ("mma", None), ("convert", None), ("epilogue", None)]
# boundaries: the first UTCHMMA belongs to the MMA role; the first F2FP to convert; the first LDTM after the last F2FP to
# the epilogue.  Walk back from each marker to the preceding unconditional BRA/EXIT (end of the previous role).
def back_to_role_start(i):
    while i > 0 and not (src[i - 1].strip().startswith("BRA ") or "EXIT" in src[i - 1]):
        i -= 1
    return i
i_mma = back_to_role_start(first("UTCHMMA") - 60)
i_conv = back_to_role_start(first("F2FP") - 120)
i_epi = back_to_role_start(next(i for i in range(last("F2FP"), len(src)) if "LDTM" in src[i]) - 150)
i_prod = back_to_role_start(first("UBLKCP") - 300)
bounds = [("prologue", 0, i_prod), ("producers", i_prod, i_mma), ("mma", i_mma, i_conv), ("convert", i_conv, i_epi),
          ("epilogue+tail", i_epi, len(src))]
tot_s = sum(int(r[ix["# Samples"]]) for r in uniq)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in uniq)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("%-14s %10s %8s %12s %8s  top stalls" % ("role", "samples", "%", "warp instr", "%"))
for name, a, b in bounds:
    s = sum(int(r[ix["# Samples"]]) for r in uniq[a:b])
    n = sum(int(r[ix["Instructions Executed"]]) for r in uniq[a:b])
    st = sorted(((sum(int(r[ix[h]] or 0) for r in uniq[a:b]), h[6:]) for h in stalls), reverse=True)[:4]
    print("%-14s %10d %7.1f%% %12d %7.1f%%  %s" % (name, s, 100.0 * s / tot_s, n, 100.0 * n / tot_i,
                                                 ", ".join("%s %.0f%%" % (h, 100.0 * v / max(s, 1)) for v, h in st)))
