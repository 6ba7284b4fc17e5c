"""Summarise an ncu report: python scripts/ncu_summary.py report.ncu-rep [kernel-substring]
Prints per-kernel headline metrics (raw page) and the hottest source lines by warp-stall samples (source page)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
        "smsp__inst_executed_op_global_st.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if filt not in name:
        continue
    print("====", name)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]:>18s} {units[i]}")
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i])) for i, h in enumerate(hdr)
          if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[i] not in ("", "0")]
    tot = sum(v for _, v in st) or 1
    print("  stalls:", ", ".join(f"{k} {v / tot * 100:.0f}%" for k, v in sorted(st, key=lambda x: -x[1])[:8]))
