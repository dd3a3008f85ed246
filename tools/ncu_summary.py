#!/usr/bin/env python
"""Key numbers of one-kernel ncu reports (raw page) + the hottest source lines (source page).
usage: tools/ncu_summary.py report.ncu-rep [n_lines]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
keys = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
for k in keys:
    if k in d: print(f"{k},{u.get(k,'')},{d[k][:100]}")
for h, v in zip(hdr, vals):
    if h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h:
        print(f"stall {h.split('stalled_')[1].split('_per')[0]},{v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
f = None; cur = None; lines = collections.OrderedDict(); iS = iI = None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": iS = r.index("# Samples"); iI = r.index("Instructions Executed"); continue
    if r[0] != "": cur = (f, int(r[0])); lines.setdefault(cur, [r[1], 0.0, 0.0, 0]); continue
    try: s = float(r[iS] or 0); i = float(r[iI] or 0)
    except Exception: continue
    L = lines[cur]; L[1] += s; L[2] += i; L[3] += 1
ts = sum(v[1] for v in lines.values()) or 1; ti = sum(v[2] for v in lines.values()) or 1
print(f"# static SASS instructions {sum(v[3] for v in lines.values())}; hottest source lines (stall samples %, executed instructions %)")
for (f, ln), v in sorted(lines.items(), key=lambda x: -x[1][1])[:nl]:
    print(f"{f}:{ln},{v[1]/ts*100:.1f},{v[2]/ti*100:.1f},{v[0].strip()[:90]}")
