"""Summarises an .ncu-rep: per kernel launch the headline raw metrics, and the hottest source lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active",
        "derived__l1tex__lsu_writeback_active_mem_lg.sum.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print("---")
    for i in idx:
        print(f"  {hdr[i]} = {r[i]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern = None; hdr = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == "Function Name": kern = r[1]; data.setdefault(kern, []); continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        try:
            data[kern].append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), int(r[0]), r[1].strip()[:100]))
        except Exception: pass
for k, d in data.items():
    ts = sum(x[0] for x in d) or 1; ti = sum(x[1] for x in d) or 1
    print(f"=== {k}: samples {ts} inst {ti}")
    for x in sorted(d, reverse=True)[:top]:
        print(f"  {x[0]/ts*100:5.1f}% smp {x[1]/ti*100:5.1f}% inst L{x[2]}: {x[3]}")
