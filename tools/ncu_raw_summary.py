#!/usr/bin/env python3
"""Key counters of one kernel from an `ncu --page raw --csv` export -> JSON on stdout."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
num = lambda k: float(d[k].replace(",", "")) if d.get(k) not in (None, "") else None
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
units = dict(zip(hdr, rows[1]))
out = {"kernel": d.get("Kernel Name"), "metrics": {k: {"value": num(k), "unit": units.get(k, "")} for k in keys}}
st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr
      if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
out["stall_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
json.dump(out, sys.stdout, indent=1)
