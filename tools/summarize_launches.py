#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share."""
import csv
import collections
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [l for l in open(path) if l.startswith('"')]
rd = csv.DictReader(rows)
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}[unit]
    k = r["Kernel Name"].split("(")[0][:60]
    tot[k] += ms
    cnt[k] += 1
total = sum(tot.values())
print("# %s\n" % title if title else "", end="")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("| %s | %d | %.2f | %.1f%% |" % (k, cnt[k], ms, 100 * ms / total))
