#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed): one CSV row per captured launch with the metrics the
roofline discussion uses.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
           "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel"] + ["%s [%s]" % (m, units[idx[m]]) for m in METRICS if m in idx])
        for r in body:
            w.writerow([r[idx["ID"]], r[idx["Kernel Name"]][:60]] + [r[idx[m]] for m in METRICS if m in idx])
    print("wrote", out, len(body), "launches")


if __name__ == "__main__":
    main()
