"""Summarise an `ncu --set full` report (.ncu-rep) as a markdown table, one column per kernel (mean over its launches).

  python tools/summarize_ncu.py gpurun_out/r1j_build.ncu-rep --title "..." > profiles/r1j_build_ncu_summary.md
  python tools/summarize_ncu.py REP --traffic k_force      # prints DRAM bytes per launch (for profiles/traffic.json)
"""
import argparse
import collections
import csv
import io
import re
import subprocess

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__inst_executed.sum",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def load(path):
    if path.endswith(".csv"):       # already exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > X_raw.csv
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    out = "\n".join(l for l in out.splitlines() if l.startswith('"'))
    # ncu 2025.2 writes an empty record after every launch and prefixes some metric names with their section
    # ("FBSP.TriageCompute.dram__..."): keep the full-width rows only (the empty ones made every later column of
    # k_force_rem read as NaN in round 1) and address metrics by their suffix
    rows = [r for r in csv.reader(io.StringIO(out)) if len(r) > 20]
    return rows[0], rows[1], rows[2:]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--title", default=None)
    ap.add_argument("--traffic", default=None, help="kernel name prefix: print mean DRAM bytes per launch")
    a = ap.parse_args()
    hdr, units, rows = load(a.report)
    col = {h: i for i, h in enumerate(hdr)}
    for i, h in enumerate(hdr):          # suffix aliases: "SECTION.metric" -> "metric" (first occurrence wins)
        if "." in h:
            for cut in range(1, h.count(".") + 1):
                tail = h.split(".", cut)[-1]
                col.setdefault(tail, i)
    groups = collections.OrderedDict()
    for r in rows:
        groups.setdefault(short(r[col["Kernel Name"]]), []).append(r)

    def mean(rs, m):
        i = col.get(m)
        if i is None:
            return None
        vals = [float(r[i].replace(",", "")) for r in rs if r[i] not in ("", "n/a", "nan")]
        if not vals:
            return None
        return sum(vals) / len(vals) * SCALE.get(units[i], 1.0)

    if a.traffic:
        for k, rs in groups.items():
            if k == a.traffic or k.startswith(a.traffic + "<"):
                print(int(mean(rs, "dram__bytes_read.sum") + mean(rs, "dram__bytes_write.sum")))
        return
    names = list(groups)
    print("# %s" % (a.title or a.report))
    print()
    print("Means over the captured launches of each kernel (`ncu --set full --clock-control none`); times in us, bytes in B.")
    print()
    print("| metric | " + " | ".join("%s (x%d)" % (k, len(groups[k])) for k in names) + " |")
    print("|---|" + "---:|" * len(names))
    for m in METRICS:
        vals = [mean(groups[k], m) for k in names]
        if all(v is None for v in vals):
            continue
        print("| %s | " % m + " | ".join("-" if v is None else ("%.4g" % v) for v in vals) + " |")
    print()
    print("| derived | " + " | ".join(names) + " |")
    print("|---|" + "---:|" * len(names))
    bw = []
    for k in names:
        t = mean(groups[k], "gpu__time_duration.sum")
        b = (mean(groups[k], "dram__bytes_read.sum") or 0) + (mean(groups[k], "dram__bytes_write.sum") or 0)
        bw.append("%.0f" % (b / (t * 1e-6) / 1e9) if t else "-")
    print("| DRAM GB/s (read+write bytes / duration) | " + " | ".join(bw) + " |")


if __name__ == "__main__":
    main()
