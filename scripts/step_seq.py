#!/usr/bin/env python
"""Launch sequence (name, ms) of the full-size resident step in an ncu launch list: from the largest seed_calls launch to the next k_compact_rows."""
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
key = "seed_calls" if any(n.startswith("seed_calls") for n, _ in seq) else "seed_smem"
big = max(v for n, v in seq if n.startswith(key))
i = [k for k, (n, v) in enumerate(seq) if n.startswith(key) and v > 0.95 * big][-1]
while i > 0 and not seq[i - 1][0].startswith("k_compact_rows") and not seq[i - 1][0].startswith("k_to_nt4") and not seq[i-1][0].startswith("k_datum"): i -= 1
tot = 0.0
while True:
    print("%-48s %8.3f" % seq[i]); tot += seq[i][1]
    if seq[i][0].startswith("k_compact_rows") or i + 1 >= len(seq): break
    i += 1
print("%-48s %8.3f" % ("sum (kernels serialised by ncu)", tot))
