#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share.  Only kernels of libhaccsr (namespace haccsr) are grouped under "ours"; torch kernels
of the snapshot generator are listed as "other".

    python tools/summarize_launches.py gpurun_out/r1_launches.csv > profiles/r1_launches_summary.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    # ncu prints libhaccsr's kernels with or without their namespace, depending on the version
    m = re.match(r"(?:void )?(?:haccsr::)?(k_\w+)(<[^>]*>)?", name)
    if m:
        return m.group(1) + (m.group(2) or ""), True
    return re.sub(r"\(.*", "", name)[:60], False


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        key, ours = short(name)
        a = agg.setdefault(key, {"n": 0, "ns": 0.0, "ours": ours, "max": 0.0})
        a["n"] += 1
        a["ns"] += ns
        a["max"] = max(a["max"], ns)
    tot_ours = sum(a["ns"] for a in agg.values() if a["ours"])
    print("# launch list summary: %s" % path)
    print()
    print("%d launches captured, %d of libhaccsr kernels; libhaccsr device time %.3f ms" % (
        len(rows), sum(a["n"] for a in agg.values() if a["ours"]), tot_ours / 1e6))
    print()
    print("| kernel | launches | total ms | share of libhaccsr time | max ms |")
    print("|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        if not a["ours"]:
            continue
        print("| %s | %d | %.3f | %.1f %% | %.3f |" % (k, a["n"], a["ns"] / 1e6, 100.0 * a["ns"] / tot_ours, a["max"] / 1e6))
    other = [(k, a) for k, a in agg.items() if not a["ours"]]
    print()
    print("other kernels (torch: snapshot generation, timing helpers): %d launches, %.3f ms" % (
        sum(a["n"] for _, a in other), sum(a["ns"] for _, a in other) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
