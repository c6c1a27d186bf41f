#!/usr/bin/env python
"""Per-kernel time of a step from an ncu launch list (gpu__time_duration.sum CSV): total per kernel name over the LAST `steps` steps
(a step ends with k_compact_rows), divided by steps.  usage: launch_summary.py launches.csv [steps]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1]))); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(r[vi].replace(",", "")); v = {"us": v / 1e3, "ns": v / 1e6, "s": v * 1e3, "ms": v}.get(r[ui], v)
    seq.append((name, v))
ends = [i for i, (n, _) in enumerate(seq) if n.startswith("k_compact_rows")]
lo = ends[-steps - 1] + 1 if len(ends) > steps else 0
agg = collections.OrderedDict(); cnt = collections.Counter()
for n, v in seq[lo:ends[-1] + 1]:
    agg[n] = agg.get(n, 0.0) + v; cnt[n] += 1
tot = sum(agg.values())
for n, v in agg.items(): print(f"{n[:60]:60s} {v / steps:8.3f} ms  x{cnt[n] / steps:g}")
print(f"{'total':60s} {tot / steps:8.3f} ms over {steps} step(s)")
