"""Summarise one kernel of an .ncu-rep: selected raw metrics + warp-stall samples per code region.
    python tools/ncu_stalls.py gpurun_out/r2b_attention.ncu-rep > profiles/r2b_attention_ncu.txt
Regions are runs of BLOCK consecutive SASS instructions, labelled by their dominant opcodes and by how often they
execute (which tells the warp role: per-step softmax code executes steps x softmax-warps times)."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
BLOCK = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                                                  text=True).stdout)))
h, u, v = raw[0], raw[1], raw[2]
want = ("gpu__time_duration.sum", "launch__registers_per_thread", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active")
print("kernel:", v[h.index("Kernel Name")] if "Kernel Name" in h else "?")
for k in want:
    if k in h:
        print(f"  {k:70s} {v[h.index(k)]:>16s} {u[h.index(k)]}")
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                                                  text=True).stdout)))
hdr = src[1]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
st = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in src[2:]:
    try:
        data.append((int(r[si]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
agg = collections.Counter()
for n, r in data:
    for i in st:
        agg[hdr[i][6:]] += int(r[i] or 0)
print(f"\nwarp-stall samples (all warps): {tot}")
print("  " + ", ".join(f"{k} {100 * c / tot:.1f}%" for k, c in agg.most_common(10)))
print(f"\nper region of {BLOCK} SASS instructions (share of samples, executions, dominant opcodes, dominant stall reasons):")
for b in range(0, len(data), BLOCK):
    seg = data[b:b + BLOCK]
    n = sum(d[0] for d in seg)
    if n * 200 < tot:
        continue

    def opcode(t):
        t = t.split()
        return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops = collections.Counter(opcode(d[1][so]) for d in seg)
    a = collections.Counter()
    for d in seg:
        for i in st:
            a[hdr[i][6:]] += int(d[1][i] or 0)
    ex = max(int(d[1][ie] or 0) for d in seg)
    print(f"  #{b:5d} {100 * n / tot:5.1f}%  exec {ex:8d} | " + ", ".join(f"{o}:{c}" for o, c in ops.most_common(4)) +
          " | " + ", ".join(f"{o}:{c}" for o, c in a.most_common(3)))
