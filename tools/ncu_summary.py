"""Summarise an .ncu-rep (ncu --set full) into one line per captured launch: duration, DRAM traffic, pipe
utilisation, occupancy.  Usage: python tools/ncu_summary.py file.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu(mufu)%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
idx = {k: hdr.index(k) for k, _ in want if k in hdr}
kn = hdr.index("Kernel Name")
print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised launches)")
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("sfm::", "").replace("<unnamed>::", "")
    parts = [f"{lab}={r[idx[k]]}{units[idx[k]] if lab in ('dur', 'dram_rd', 'dram_wr') else ''}" for k, lab in want if k in idx]
    print(f"{name[:46]:46s} " + " ".join(parts))
