"""ncu .ncu-rep -> one markdown row per captured kernel with the metrics the roofline notes quote."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
def col(name):
    return hdr.index(name) if name in hdr else None
cols = {"name": col("Kernel Name"), "t": col("gpu__time_duration.sum"), "inst": col("smsp__inst_executed.sum"),
        "issue": col("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps": col("sm__warps_active.avg.per_cycle_active"),
        "fma": col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), "alu": col("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "rd": col("dram__bytes_read.sum"), "wr": col("dram__bytes_write.sum"), "dram": col("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor": col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "l2": col("lts__t_sector_hit_rate.pct"),
        "regs": col("launch__registers_per_thread"), "grid": col("launch__grid_size"), "lsu": col("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
        "longsb": col("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio")}
units = rows[1]
print("| kernel | time | warp inst | issue % | warps/SM | FMA % | ALU % | tensor % | DRAM rd | DRAM wr | DRAM % | L2 hit % | regs | long_sb |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in rows[2:]:
    g = lambda k: (r[cols[k]] if cols[k] is not None else "")
    u = lambda k: (units[cols[k]] if cols[k] is not None else "")
    f = lambda k: (f"{float(g(k).replace(',', '')):.1f}" if g(k) not in ("", "n/a") else "")
    print(f"| `{g('name')[:60]}` | {g('t')} {u('t')} | {float(g('inst').replace(',',''))/1e6:.1f} M | {f('issue')} | {f('warps')} | {f('fma')} | {f('alu')} | {f('tensor')} | {g('rd')} {u('rd')} | {g('wr')} {u('wr')} | {f('dram')} | {f('l2')} | {g('regs')} | {f('longsb')} |")
