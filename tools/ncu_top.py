"""Summarise an ncu report: headline metrics + the most-sampled SASS instructions with their stall reasons."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, r = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for i, k in enumerate(h):
    if k in keys or ("issue_stalled" in k and "per_issue_active" in k): print(f"{k:95s} {r[i]} {u[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]; data = rows[2:]
isrc, ie, isamp, ia = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Address")
stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
ts = sum(int(x[isamp] or 0) for x in data); ti = sum(int(x[ie] or 0) for x in data)
print("total samples", ts, "total warp-instructions", ti)
op = collections.Counter()
for x in data:
    toks = x[isrc].split(); o = toks[1] if toks[0].startswith("@") else toks[0]
    op[o.split(".")[0]] += int(x[ie] or 0)
print("opcode mix:", ", ".join(f"{o} {100*n/ti:.1f}%" for o, n in op.most_common(16)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:ntop]
for i in sorted(top):
    x = data[i]
    st = sorted(((h[j][6:], int(x[j] or 0)) for j in stalls if int(x[j] or 0) > 0), key=lambda p: -p[1])[:3]
    print(i, x[ia][-5:], f"exec {int(x[ie] or 0):9d}", f"{100*int(x[isamp])/ts:5.2f}%", x[isrc][:64], st)
