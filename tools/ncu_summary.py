"""Summarise an .ncu-rep (raw page CSV) into the handful of counters DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
pat = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
                       "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "registers_per_thread", "shared_mem_per_block",
                       "occupancy_limit", "sm__inst_executed_pipe_fma", "pipe_fma_cycles_active", "pipe_tensor", "issue_active",
                       "bank_conflicts_pipe_lsu_mem_shared.sum", "warps_issue_stalled", "grid_size", "waves_per_multiprocessor",
                       "lts__t_sector_hit_rate", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "achieved_occupancy",
                       "sm__sass_thread_inst_executed_op_ffma_pred_on.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=" * 100)
    print(r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p in h for p in pat):
            v = r[i]
            if v in ("", "0", "n/a"): continue
            print(f"  {h:95s} {v:>18s} {units[i]}")
