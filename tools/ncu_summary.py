"""Markdown summary of an ncu --set full report (one section per captured kernel).  usage: ncu_summary.py report.ncu-rep title"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# {title}\n")
for r in rows[2:]:
    print(f"## {r[h.index('Kernel Name')]}")
    for k in keys:
        if k in h:
            print(f"- {k} = {r[h.index(k)]} {rows[1][h.index(k)]}")
    st = []
    for i, c in enumerate(h):
        if "issue_stalled" in c and "not_issued" not in c and c.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[i]), c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("- top stalls (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
    print()
