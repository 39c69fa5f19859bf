#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/<tag>_launches_c2.csv [profiles/traffic_c2.json] > profiles/<tag>_launches_c2.md
  python profiles/summarize.py full     gpurun_out/<tag>_prof.ncu-rep     > profiles/<tag>_prof_summary.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path, traffic_json=None):
    """Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]`
    launch list; with DRAM metrics also the per-iteration DRAM traffic of the ALS solve launches (last iteration)."""
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, im, iv, iu, iid = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            v = v / 1e6 if r[iu] in ("nsecond", "ns") else (v / 1e3 if r[iu] in ("usecond", "us") else v)  # -> ms
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(r[iu], 1.0)
        name = r[ik].split("(")[0].replace("void ", "").replace("bfl::", "").replace("tc::", "")
        per.setdefault(r[iid], {"name": name})[r[im]] = v
    ours_prefixes = ("als_", "gram_", "fast_", "bpr_", "warp_", "sgd_", "probe_", "tc_")
    agg = OrderedDict()
    for d in per.values():
        name = d["name"] if d["name"].startswith(ours_prefixes) else "torch (workload generation)"
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "torch" not in k)
    print("| kernel | launches | total ms | share of our kernels | DRAM GB (rd+wr) |\n|---|---|---|---|---|")
    for k, (n, v, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %s | %.1f |" % (k, n, v, "%.1f %%" % (100 * v / ours) if "torch" not in k else "-", b / 1e9))
    print("\ntotal %.1f ms, ours %.1f ms (ncu: cold-cache, serialised launches -- compare shares, not absolutes)" % (tot, ours))
    # last iteration: the solve launches after the last two Gram reductions (user side, then item side)
    seq = list(per.values())
    gi = [i for i, d in enumerate(seq) if d["name"].startswith("gram_reduce")]
    if len(gi) >= 2 and "dram__bytes_read.sum" in seq[gi[-1]]:
        def solve(lo, hi):
            ds = [d for d in seq[lo:hi] if d["name"].startswith("als_")]
            return (sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in ds),
                    sum(d["gpu__time_duration.sum"] for d in ds), len(ds))
        ub, ut, un = solve(gi[-2], gi[-1])
        ib, it, inn = solve(gi[-1], len(seq))
        print("\nlast iteration: user-side solve %d launches %.1f ms %.1f GB DRAM; item-side solve %d launches %.1f ms %.1f GB DRAM"
              % (un, ut, ub / 1e9, inn, it, ib / 1e9))
        if traffic_json:
            import json
            json.dump({"user_pass": ub, "item_pass": ib, "unit": "byte",
                       "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, last iteration of `%s`" % path},
                      open(traffic_json, "w"), indent=1)


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
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2])
