"""Summarise an .ncu-rep: headline metrics + instructions per warp-step by source line.
usage: python scripts/ncu_summary.py REPORT.ncu-rep WARPS STEPS [top]"""
import csv, subprocess, sys, io
rep, warps, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
for k in keys:
    if k in m: print(f"{k:75s} {m[k]:>18s} {u[k]}")
for k in hdr:
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(m[k] or 0) > 0.05:
        print(f"{k:75s} {m[k]:>18s}")
print(f"{'warp instructions per warp-step':75s} {float(m['smsp__inst_executed.sum'])/(warps*steps):18.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
def num(x):
    return int(x) if x and x.strip().lstrip('-').isdigit() else 0
agg, h, cur = [], None, ""
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": h = r; continue
    if h and r[0].isdigit():
        d = dict(zip(h, r))
        agg.append((cur.split("/")[-1], int(r[0]), r[1].strip()[:84], num(d["Instructions Executed"]), num(d["# Samples"]), num(d["Thread Instructions Executed"])))
tot = sum(a[3] for a in agg) or 1; tots = sum(a[4] for a in agg) or 1
print(f"\n{'line':>28s} {'inst/warp-step':>14s} {'%inst':>6s} {'%samples':>8s} {'thr/inst':>8s}")
for a in sorted(agg, key=lambda x: -x[3])[:top]:
    print(f"{a[0][:20]:>20s}:{a[1]:<6d} {a[3]/(warps*steps):14.1f} {100*a[3]/tot:6.1f} {100*a[4]/tots:8.1f} {a[5]/max(1,a[3]):8.1f}  {a[2]}")
