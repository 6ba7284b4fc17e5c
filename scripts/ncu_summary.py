"""Summary of an ncu report for profiles/: per kernel launch the metrics DESIGN.md quotes (duration, registers, warps,
instructions, DRAM bytes, shared-memory conflicts, issue %, stall reasons).   python scripts/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
ik = h.index("Kernel Name")
for r in rows[2:]:
    if len(r) < len(h):
        continue
    print("====", r[ik])
    for w in want:
        if w in h:
            print(f"  {w:75s} {r[h.index(w)]:>16s} {rows[1][h.index(w)]}")
    st = [(float(r[i] or 0), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i, n in enumerate(h)
          if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n]
    tot = sum(v for v, _ in st) or 1
    print("  stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:8]))
