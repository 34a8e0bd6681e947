"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python profiles/summarize_launches.py gpurun_out/r1_launches_bench.csv profiles/r1_launches_bench.md "title"
"""
import csv
import sys
from collections import defaultdict

src, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1e-6)
    tot[r[ix["Kernel Name"]]] += v
    cnt[r[ix["Kernel Name"]]] += 1
total = sum(tot.values())
lines = ["# %s" % title, "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares). Raw list: `%s`." % src.split("/")[-1], "",
         "| launches | total ms | share | kernel |", "|---|---|---|---|"]
for k in sorted(tot, key=tot.get, reverse=True):
    lines.append("| %d | %.3f | %.2f %% | `%s` |" % (cnt[k], tot[k], 100 * tot[k] / total, k[:100]))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, "total ms %.1f" % total)
