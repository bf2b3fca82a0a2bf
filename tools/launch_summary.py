"""ncu launch list (--metrics gpu__time_duration.sum --csv) → per-kernel totals and shares.  usage: launch_summary.py <csv> [header text]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v_us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    n, t = tot.get(r[ki], (0, 0.0))
    tot[r[ki]] = (n + 1, t + v_us)
total = sum(t for _, t in tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} n={n:4d} total_ms={t / 1e3:9.3f} mean_us={t / n:9.1f} share={100 * t / total:5.1f}%")
