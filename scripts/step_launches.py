#!/usr/bin/env python
"""Per-launch times of ONE resident 1M-read step, from an ncu launch list (gpu__time_duration.sum CSV)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    v = float(r[vi].replace(",", "")); v = {"us": v / 1e3, "ns": v / 1e6, "s": v * 1e3, "ms": v}.get(r[ui], v)
    seq.append((name, v))
big = max(v for n, v in seq if n.startswith("seed_smem"))
for i, (n, v) in enumerate(seq):
    if n.startswith("seed_smem") and v > 0.9 * big:
        j = i
        while j < len(seq) and not (j > i and seq[j][0].startswith("seed_smem")) and not seq[j][0].startswith("k_to_nt4"):
            print(f"{seq[j][0][:50]:50s} {seq[j][1]:8.3f} ms"); j += 1
        break
