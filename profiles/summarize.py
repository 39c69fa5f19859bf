#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/<tag>_launches_c2.csv  > profiles/<tag>_launches_c2.md
  python profiles/summarize.py full     gpurun_out/<tag>_prof.ncu-rep     > profiles/<tag>_prof_summary.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e6 if r[iu] in ("nsecond", "ns") else (v / 1e3 if r[iu] in ("usecond", "us") else v)  # -> ms
        name = r[ik].split("(")[0].replace("void ", "").replace("bfl::", "")
        ours_prefixes = ("als_", "gram_", "fast_", "bpr_", "warp_", "sgd_", "probe_")
        name = name if name.startswith(ours_prefixes) else "torch (workload generation)"
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    ours = sum(v for k, (_, v) in agg.items() if "torch" not in k)
    print("| kernel | launches | total ms | share of our kernels |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %s |" % (k, n, v, "%.1f %%" % (100 * v / ours) if "torch" not in k else "-"))
    print("\ntotal %.1f ms, ours %.1f ms (ncu: cold-cache, serialised launches -- compare shares, not absolutes)" % (tot, ours))


METRICS = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "GB rd"), ("dram__bytes_write.sum", "GB wr"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
           ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC/SM"),
           ("smsp__inst_executed.sum", "Ginst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
           ("launch__registers_per_thread", "regs")]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle", "not_selected"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, m):
        if m not in ix or r[ix[m]] in ("", "n/a"):
            return float("nan")
        v, u = float(r[ix[m]].replace(",", "")), units[ix[m]]
        scale = {"nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3,
                 "Gbyte": 1.0, "Tbyte": 1e3}.get(u, 1.0)
        return v * scale
    print("| kernel | " + " | ".join(n for _, n in METRICS) + " | stalls per issue (" + ", ".join(STALLS) + ") |")
    print("|---|" + "---|" * (len(METRICS) + 1))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for m, n in METRICS:
            v = val(r, m)
            vals.append("%.3g" % (v / 1e9 if n == "Ginst" else v))
        st = ["%.2f" % val(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s) for s in STALLS]
        print("| `%s` | %s | %s |" % (name, " | ".join(vals), ", ".join(st)))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
