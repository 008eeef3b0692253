"""One line per kernel instance of an .ncu-rep (ncu --set full): duration, tensor pipe, issue slots, warps, DRAM bytes,
shared-memory pipe, L2 hit rate.  python tools/ncu_brief.py file.ncu-rep [max_per_kernel]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 2
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
cols = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pipe%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst")]
idx = [(hdr.index(k), n) for k, n in cols if k in hdr]
seen = {}
print(f"# {rep}: ncu --set full --clock-control none; one line per captured launch (at most {cap} per kernel)")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("(")[0].replace("void ", "")[:44]
    seen[short] = seen.get(short, 0) + 1
    if seen[short] > cap:
        continue
    print(f"{short:46s} grid {r[hdr.index('Grid Size')]:>16s} " + " ".join(f"{n}={r[i]}{units[i] if n.startswith('dram_') else ''}" for i, n in idx))
