#!/usr/bin/env python3
"""Static SASS of one kernel of a binary, cut at RET (the out-of-line device functions): instruction counts per segment.
Usage: sass_segments.py <binary> <kernel-substring> [min_wide=150]"""
import collections, re, subprocess, sys
binary, kern = sys.argv[1], sys.argv[2]
min_wide = int(sys.argv[3]) if len(sys.argv) > 3 else 150
sass = subprocess.check_output(["cuobjdump", "-sass", binary], text=True)
cur, on, segs = [], False, []
for line in sass.splitlines():
    if "Function :" in line:
        on = kern in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
    if not m:
        continue
    ins = m.group(1).strip()
    toks = ins.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    cur.append(op.replace(".U32", "").replace(".LUT", ""))
    if op.startswith("RET"):
        segs.append(cur)
        cur = []
if cur:
    segs.append(cur)
print("%s: %d instructions in %d segments" % (kern, sum(len(s) for s in segs), len(segs)))
for k, s in enumerate(segs):
    c = collections.Counter(s)
    wide = c["IMAD.WIDE"] + c["IMAD.WIDE.X"]
    if wide < min_wide or len(s) > 1500:
        continue
    print("  seg %2d: %4d instrs, wide %3d | %s" % (k, len(s), wide, ", ".join("%s %d" % kv for kv in c.most_common(12))))
