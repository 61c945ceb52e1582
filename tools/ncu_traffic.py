#!/usr/bin/env python
"""Per-kernel DRAM bytes, warp instructions and duration from an `ncu --set full` report (read here, no GPU), as JSON:
the file bench.py reads for `roofline.traffic` and the voxelizer's issue roofline.
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/r02_traffic_c2.json
The LAST captured launch of every kernel counts (steady state)."""
import csv
import io
import json
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, m, scale_by_unit=True):
        v = num(r[ix[m]])
        u = units[ix[m]]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "inst": 1.0}.get(u, 1.0)
        return None if v is None else v * mult

    res = {}
    for r in body:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip()
        res[name] = {
            "kernel_full_name": r[ix["Kernel Name"]][:80],
            "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
            "warp_instructions": val(r, "smsp__inst_executed.sum"), "duration_us": val(r, "gpu__time_duration.sum"),
            "registers": val(r, "launch__registers_per_thread"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "threads_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        }
    res["_source"] = {"report": rep, "how": "ncu --set full --clock-control none, C2 step (tools/profile_step.py c2), last captured launch of each kernel"}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out, list(k for k in res if not k.startswith("_")))


if __name__ == "__main__":
    main()
