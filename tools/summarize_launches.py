#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum + dram__bytes_{read,write}.sum, --csv) per pipeline stage.

usage: tools/summarize_launches.py launches.csv [launches_per_step]
Prints, for the LAST complete step in the list, per stage: launches, time (us), DRAM read / write bytes; and the
share of every stage.  ncu serialises the launches and flushes caches between them, so the shares are meaningful,
the absolute times are not (B200_PROFILING.md)."""
import csv, json, sys, collections

def stage_of(name, grid):
    if "analysis_tma" in name: return "analysis_l1"
    if "analysis_kernel" in name: return "analysis_deep_or_l1"
    if "hist_kernel" in name: return "histogram"
    if "otsu" in name: return "otsu"
    if "filter_rows" in name or "notch_umma" in name or "dense" in name: return "row_filter"
    if "synth_kernel<1" in name or "synth_kernel<true" in name: return "final_synthesis_epilogue"
    if "synth_kernel" in name: return "synthesis_deep"
    return "other"

rows = collections.OrderedDict()
with open(sys.argv[1]) as fp:
    rd = csv.reader(l for l in fp if l.startswith('"'))
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr): continue
        k = int(r[ix["ID"]])
        d = rows.setdefault(k, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"].replace(",", "")] if False else r[ix["Metric Value"]].replace(",", ""))
launches = list(rows.values())
# a step starts at every level-1 analysis launch
# (only steps over whole chunks: the end-to-end part of the bench runs 4-plane sub-chunks)
zmax = max(int(d["grid"].strip("()").split(",")[-1]) for d in launches if "analysis_tma" in d["name"])
starts = [i for i, d in enumerate(launches) if "analysis_tma" in d["name"] and d["grid"].strip("()").split(",")[-1].strip() == str(zmax)]
per = int(sys.argv[2]) if len(sys.argv) > 2 else (starts[1] - starts[0] if len(starts) > 1 else len(launches))
s0 = starts[1] if len(starts) > 2 else starts[0]
step = launches[s0:s0 + per]
agg = collections.OrderedDict()
first_rows = True
for d in step:
    st = stage_of(d["name"], d["grid"])
    if st == "row_filter" and first_rows:
        st, first_rows = "row_filter_level1", False
    a = agg.setdefault(st, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    a["launches"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    a["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
    a["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
tot_us = sum(a["us"] for a in agg.values())
tot_b = sum(a["dram_read"] + a["dram_write"] for a in agg.values())
out = {"launches_in_step": len(step), "sum_us": tot_us, "dram_bytes_per_step": tot_b, "stages": {}}
for k, a in agg.items():
    a["share_of_time"] = a["us"] / tot_us
    a["dram_bytes"] = a["dram_read"] + a["dram_write"]
    out["stages"][k] = a
print(json.dumps(out, indent=1))
