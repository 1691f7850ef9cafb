"""dev: key numbers of one .ncu-rep (first kernel): time, issue utilisation, pipes, stalls, occupancy, top source lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.3):
        print(f"{h:90s} {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; out = []
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "File Path": cur = row[1].split('/')[-1]; continue
    if cur and len(row) > 8 and row[0].isdigit():
        try: ie = int(row[7])
        except ValueError: continue
        out.append((ie, cur, int(row[0]), row[1].strip()[:120], row[6]))
tot = sum(o[0] for o in out)
print("total warp instr", tot)
for ie, f, l, s, samp in sorted(out, reverse=True)[:top]:
    print(f"{100 * ie / tot:5.1f}% {ie:9d} samp={samp:>5} {f}:{l}  {s}")
