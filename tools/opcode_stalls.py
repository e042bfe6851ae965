#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` export by SASS opcode: share of executed warp instructions, share of stall
samples and the top stall reasons per opcode.  Usage: opcode_stalls.py source.csv [title]"""
import collections
import csv
import sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ex = collections.Counter()
smp = collections.Counter()
why = collections.defaultdict(collections.Counter)
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    if not src:
        continue
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    parts = op.split(".")
    key = ".".join(parts[:2]) if parts[0] in ("IMAD", "IADD3", "LDL", "STL", "LD", "ST") and len(parts) > 1 and parts[1] in (
        "WIDE", "X", "HI", "MOV", "IADD", "SHL", "128", "64", "E") else parts[0]
    if op.startswith("IMAD.WIDE"):
        key = "IMAD.WIDE.X" if ".X" in op else "IMAD.WIDE"
    ex[key] += int(r[col["Instructions Executed"]] or 0)
    smp[key] += int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    for s in stall_cols:
        v = int(r[col[s]] or 0)
        if v:
            why[key][s[len("stall_"):]] += v
te, ts = sum(ex.values()), sum(smp.values())
print("# %s\n" % title)
print("executed warp instructions: %.3f G, stall samples: %d\n" % (te / 1e9, ts))
print("| opcode | % of executed instructions | % of stall samples | top stall reasons (share of the opcode's samples) |")
print("|---|---|---|---|")
for k, v in ex.most_common(16):
    tot = sum(why[k].values()) or 1
    top = ", ".join("%s %d%%" % (n, round(100 * c / tot)) for n, c in why[k].most_common(3))
    print("| %s | %.2f | %.2f | %s |" % (k, 100 * v / te, 100 * smp[k] / max(ts, 1), top))
