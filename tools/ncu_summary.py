"""Tracked summaries of an .ncu-rep (run here, no GPU): python tools/ncu_summary.py <report> <profiles/prefix>
writes <prefix>_raw.csv (selected raw metrics per profiled launch) and <prefix>_stalls.txt (warp-stall sampling)."""
import csv, subprocess, sys

rep, prefix = sys.argv[1:3]
WANT = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(n) for n in WANT if n in hdr]
with open(prefix + "_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i][:90] for i in idx])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
name = rows[0][1]; h = {n: i for i, n in enumerate(rows[1])}
stalls = [n for n in rows[1] if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}; total = n = 0
for r in rows[2:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    n += 1; total += int(r[h["# Samples"]])
    for s in stalls: tot[s] += int(r[h[s]])
with open(prefix + "_stalls.txt", "w") as f:
    f.write(f"kernel: {name}\nSASS instructions: {n}   warp-stall samples: {total}\n")
    for s, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v: f.write(f"  {s:28s}{v:8d} {100*v/max(total,1):5.1f}%\n")
print(open(prefix + "_raw.csv").read()); print(open(prefix + "_stalls.txt").read())
