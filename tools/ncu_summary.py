#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the short text summary kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "title line" "command line" > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {title}")
    print(f"# command: {command}")
    print("# (profiler run: launches are serialised and cold, and the outputs of a launch are still in the 126 MB L2 when it"
          " ends, so dram write bytes under-count; timing numbers come from bench.py, not from here)")
    for n, r in enumerate(data):
        print(f"## launch {n}")
        print("Kernel Name =", r[hdr.index("Kernel Name")])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k} = {r[i]} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"{h} = {v:.6f}")


if __name__ == "__main__":
    main()
