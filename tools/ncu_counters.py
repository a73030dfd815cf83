#!/usr/bin/env python3
"""Turns the ncu launch list of `python bench.py --steps K --warmup W` (--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,...) into the per-launch averages
bench.py quotes in its roofline object (profiles/r2_kernel_counters.json).
usage: python tools/ncu_counters.py launches.csv reads read_len ref_bases "source text" [useful_lanes_full useful_lanes_first] > profiles/r2_kernel_counters.json"""
import collections
import csv
import json
import sys

path, reads, read_len, ref_bases, source = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
useful = [float(x) for x in sys.argv[6:8]] if len(sys.argv) >= 8 else [None, None]
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
col = {h: i for i, h in enumerate(rows[hi])}
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[hi + 1:]:
    if len(r) < len(col):
        continue
    name = r[col["Kernel Name"]]
    key = "full" if ("xm_align_kernel<0>" in name or "(bool)0" in name) else "first_pass" if ("xm_align_kernel<1>" in name or "(bool)1" in name) else None
    if key is None:
        continue
    try:
        agg[key][r[col["Metric Name"]]].append(float(r[col["Metric Value"]].replace(",", "")))
    except ValueError:
        pass
out = dict(workload=dict(reads=reads, read_len=read_len, paired=False, ref_bases=ref_bases), source=source)
for i, key in enumerate(("full", "first_pass")):
    m = agg[key]
    avg = lambda k: (sum(m[k]) / len(m[k])) if m.get(k) else None
    out[key] = dict(launches=len(m.get("gpu__time_duration.sum", [])), ncu_ms_per_launch=(avg("gpu__time_duration.sum") or 0) / 1e6,
                    dram_bytes_read=avg("dram__bytes_read.sum"), dram_bytes_write=avg("dram__bytes_write.sum"),
                    warp_instructions=avg("smsp__inst_executed.sum"), active_lanes_per_instruction=avg("smsp__thread_inst_executed_per_inst_executed.ratio"),
                    issue_active_pct=avg("smsp__issue_active.avg.pct_of_peak_sustained_active"), fp64_pipe_warp_instructions=avg("sm__inst_executed_pipe_fp64.sum"),
                    local_load_warp_instructions=avg("smsp__inst_executed_op_local_ld.sum"), local_store_warp_instructions=avg("smsp__inst_executed_op_local_st.sum"),
                    useful_lanes_per_instruction=useful[i])
print(json.dumps(out, indent=1))
