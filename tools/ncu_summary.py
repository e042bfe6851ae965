#!/usr/bin/env python3
"""Builds profiles/<tag>_ncu_summary.json (the ncu-derived numbers bench.py's roofline blocks quote) and per-kernel
opcode/stall tables from the exports tools/gpu_profile.sh leaves in gpurun_out/:
    <tag>_prof_raw.csv            `ncu --page raw --csv` of one launch of every hot kernel at 2^20 items
    <tag>_src_<kernel>.csv.gz     `ncu --page source --csv` per kernel
Usage: tools/ncu_summary.py <tag> [n_items=1048576]"""
import collections
import csv
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
n_items = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
G = os.path.join(ROOT, "gpurun_out")
csv.field_size_limit(1 << 30)

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def num(v):
    try:
        x = float(v.replace(",", ""))
        return None if x != x else x  # ncu prints -nan for a metric it could not collect
    except Exception:
        return None


SCALE = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def dram_bytes(d):
    """read + write bytes of one launch; ncu picks a unit PER METRIC and per session (d carries them as <metric>@unit)"""
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v = num(d.get(k, ""))
        if v is None:
            return None
        tot += v * SCALE[d.get(k + "@unit", "byte")]
    return tot


def short(name):
    base = name.replace("(int)", "").split("(")[0].replace("void ", "").replace("sylow_kernels::", "").strip()
    return base.replace(" ", "")  # "void k_glued<(int)1, (int)3>(...)" -> "k_glued<1,3>"


def opcode_table(path):
    with gzip.open(path, "rt") as f:
        rows = list(csv.reader(f))
    tables = []
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            col = {h: j for j, h in enumerate(hdr)}
            stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            ex, smp, why = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
            j = i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                r = rows[j]
                j += 1
                if len(r) < len(hdr) or not r[col["Source"]].strip():
                    continue
                toks = r[col["Source"]].split()
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                parts = op.split(".")
                key = ".".join(parts[:2]) if parts[0] in ("IMAD", "IADD3", "LDL", "STL", "LD", "ST") and len(parts) > 1 and \
                    parts[1] in ("WIDE", "X", "HI", "MOV", "IADD", "SHL", "128", "64", "E") else parts[0]
                if op.startswith("IMAD.WIDE"):
                    key = "IMAD.WIDE.X" if ".X" in op else "IMAD.WIDE"
                ex[key] += int(r[col["Instructions Executed"]] or 0)
                smp[key] += int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
                for s in stall_cols:
                    v = int(r[col[s]] or 0)
                    if v:
                        why[key][s[len("stall_"):]] += v
            tables.append((name, ex, smp, why))
            i = j
        else:
            i += 1
    return tables


def main():
    summary = {}
    launches = collections.defaultdict(list)
    main = os.path.join(G, "%s_prof_raw.csv" % tag)
    rows = list(csv.reader(open(main))) if os.path.exists(main) else [[], []]
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name") if hdr else 0
    def with_units(h, u, r):
        d = dict(zip(h, r))
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in h:
                d[k + "@unit"] = u[h.index(k)]
        return d

    for r in rows[2:]:
        launches[short(r[ki])].append(with_units(hdr, units, r))
    # per-kernel sessions (tools/gpu_profile_one.sh): in a session with many kernels ncu sometimes collects only the
    # source-level passes for a kernel (6 instead of 39 passes, hardware counters "-nan"); a launch with hardware
    # counters from its own session replaces the incomplete one
    for fn in sorted(os.listdir(G)):
        if fn.startswith(tag + "_raw_") and fn.endswith(".csv"):
            rr = list(csv.reader(open(os.path.join(G, fn))))
            for r in rr[2:]:
                d = with_units(rr[0], rr[1], r)
                if num(d.get("smsp__inst_executed.sum", "")) is None:
                    continue
                k = short(d["Kernel Name"])
                same = [x for x in launches[k] if x["launch__grid_size"] == d["launch__grid_size"]]
                for x in same:
                    launches[k].remove(x)
                launches[k].append(d)
    md = []
    for kname, ls in launches.items():
        # the kernel's main launch (largest grid); a low-occupancy tail launch of the same kernel is listed beside it
        ls.sort(key=lambda d: -(num(d["launch__grid_size"]) or 0) * (num(d["launch__block_size"]) or 0))
        d = ls[0]
        items = n_items
        threads = (num(d["launch__grid_size"]) or 0) * (num(d["launch__block_size"]) or 0)
        if kname.startswith("k_glued<1,3>") or kname.startswith("k_glued<4,0>") or kname == "k_check_products":
            items = n_items // 4
        elif 0 < threads < n_items:
            items = int(threads)  # the kernel's main launch covers the whole waves only (a tail launch does the rest)
        m = {k: num(d.get(k, "")) for k in KEYS}
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: num(d[h]) for h in d
                  if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and num(d[h]) is not None}
        entry = {"n": items, "ms": m["gpu__time_duration.sum"], "grid": m["launch__grid_size"], "block": m["launch__block_size"],
                 "registers": m["launch__registers_per_thread"],
                 "fmaheavy_pct": m["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"],
                 "alu_pct": m["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"],
                 "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                 "inst_executed": m["smsp__inst_executed.sum"],
                 "dram_bytes": dram_bytes(d),
                 "local_ld_hit_pct": m["l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"],
                 "local_st_hit_pct": m["l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct"],
                 "stall_per_issue": dict(sorted(stalls.items(), key=lambda kv: -(kv[1] or 0))[:8]),
                 "source": "profiles/%s_ncu_summary.json <- ncu --set full, one launch at %d items (tools/gpu_profile.sh)" % (tag, items)}
        if len(ls) > 1:
            entry["tail_launch"] = {"ms": num(ls[1]["gpu__time_duration.sum"]), "grid": num(ls[1]["launch__grid_size"]),
                                    "block": num(ls[1]["launch__block_size"])}
        summary[kname] = entry
    # opcode shares from the source pages
    for fn in sorted(os.listdir(G)):
        if not (fn.startswith(tag + "_src_") and fn.endswith(".csv.gz")):
            continue
        for name, ex, smp, why in opcode_table(os.path.join(G, fn)):
            kname = short(name)
            te, ts = sum(ex.values()), sum(smp.values())
            if not te:
                continue
            wide = ex["IMAD.WIDE"] + ex["IMAD.WIDE.X"]
            if kname in summary and (summary[kname].get("imad_wide_share") is None or te > summary[kname].get("_te", 0)):
                summary[kname]["imad_wide_share"] = wide / te
                summary[kname]["_te"] = te
                if not summary[kname].get("inst_executed"):
                    summary[kname]["inst_executed"] = float(te)
                summary[kname]["opcode_pct"] = {k: round(100.0 * v / te, 2) for k, v in ex.most_common(12)}
            md.append("## %s: ncu source page aggregated by opcode\n\nexecuted warp instructions: %.3f G, stall samples: %d\n" % (kname, te / 1e9, ts))
            md.append("| opcode | % of executed instructions | % of stall samples | top stall reasons (share of the opcode's samples) |\n|---|---|---|---|")
            for k, v in ex.most_common(16):
                tot = sum(why[k].values()) or 1
                top = ", ".join("%s %d%%" % (n, round(100 * c / tot)) for n, c in why[k].most_common(3))
                md.append("| %s | %.2f | %.2f | %s |" % (k, 100 * v / te, 100 * smp[k] / max(ts, 1), top))
            md.append("")
    for e in summary.values():
        e.pop("_te", None)
    out = os.path.join(ROOT, "profiles", "%s_ncu_summary.json" % tag)
    json.dump(summary, open(out, "w"), indent=1, sort_keys=True)
    open(os.path.join(ROOT, "profiles", "%s_opcode_stalls.md" % tag), "w").write("\n".join(md) + "\n")
    for k, e in summary.items():
        print("%-18s n=%-8d %8.3f ms  fmaheavy %s  issue %s  inst %.3g  wide-share %s  dram %s" % (
            k, e["n"], e["ms"] or -1, e["fmaheavy_pct"], e["issue_active_pct"], e["inst_executed"] or 0,
            None if e.get("imad_wide_share") is None else round(e["imad_wide_share"], 4),
            None if e["dram_bytes"] is None else "%.3g GB" % (e["dram_bytes"] / 1e9)))


if __name__ == "__main__":
    main()
