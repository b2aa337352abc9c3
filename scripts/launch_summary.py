"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / total / share.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [skip_first_n_launches]"""
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
    rows.append((r[ix["Kernel Name"]], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = {}
for name, us in rows:
    name = re.sub(r"\(.*", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, total {tot / 1e3:.3f} ms")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  n={n:4d}  avg={us / n:8.1f} us  {name[:100]}")
