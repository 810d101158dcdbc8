#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` of one kernel launch into a short markdown table of the metrics
that matter for an FMA-pipe- or HBM-bound kernel.
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv
    python tools/summarize_ncu.py /tmp/x.csv "title" [pairs_in_launch] > profiles/x_summary.md
"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_blocks", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    path, title = sys.argv[1], sys.argv[2]
    pairs = float(sys.argv[3]) if len(sys.argv) > 3 else None
    row = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("# %s\n" % title)
    print("| metric | unit | value |\n|---|---|---|")
    for w in WANT:
        if w in d:
            print("| %s | %s | %s |" % (w, d[w][0], d[w][1]))
    if pairs:
        t = float(d["gpu__time_duration.sum"][1].replace(",", "")) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[d["gpu__time_duration.sum"][0]]
        inst = float(d["smsp__inst_executed.sum"][1].replace(",", ""))
        print("\nDerived: %.3e pairs in this launch -> %.3f T pairs/s under the profiler = %.1f TFLOP/s at 30 flop/pair = "
              "%.1f %% of the 74.45 TFLOP/s FP32 peak; %.1f warp instructions per warp-pair (32 sinks x 1 source)." % (
                  pairs, pairs / t / 1e12, 30 * pairs / t / 1e12, 100 * 30 * pairs / t / 74.45e12, inst / (pairs / 32)))


if __name__ == "__main__":
    main()
