"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
of one UNet forward: per kernel name count / time / DRAM bytes, and the per-launch DRAM traffic of the dominant
kernel (gemm_tcgen05_kernel) that bench.py reports as roofline.traffic.
usage: python scripts/traffic_summary.py gpurun_out/launches_dram.csv <skip_first_n_launches> [out.json]"""
import csv
import json
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
per = {}
order = []
for r in rd:
    if len(r) != len(hdr):
        continue
    kid = r[ix["ID"]]
    if kid not in per:
        per[kid] = dict(name=re.sub(r"\(.*", "", r[ix["Kernel Name"]]), us=0.0, rd=0.0, wr=0.0)
        order.append(kid)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        per[kid]["us"] = v / 1e3 if unit in ("nsecond", "ns") else (v if unit in ("usecond", "us") else v * 1e3)
    else:
        mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[kid]["rd" if "read" in m else "wr"] = v * mul
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [per[k] for k in order][skip:]
agg = {}
for r in rows:
    a = agg.setdefault(r["name"], dict(n=0, us=0.0, rd=0.0, wr=0.0))
    a["n"] += 1
    a["us"] += r["us"]
    a["rd"] += r["rd"]
    a["wr"] += r["wr"]
tot = sum(a["us"] for a in agg.values())
print(f"{len(rows)} launches, total {tot / 1e3:.3f} ms (cold-cache, serialised: compare shares)")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{a['us'] / 1e3:8.3f} ms {100 * a['us'] / tot:5.1f}%  n={a['n']:4d}  avg={a['us'] / a['n']:7.1f} us  "
          f"dram rd {a['rd'] / 1e6:8.1f} MB wr {a['wr'] / 1e6:8.1f} MB  {name[:70]}")
g = [a for n, a in agg.items() if "gemm_tcgen05" in n]
if g:
    n = sum(a["n"] for a in g)
    out = dict(kernel="gemm_tcgen05_kernel", launches=n, dram_bytes_per_launch=(sum(a["rd"] + a["wr"] for a in g)) / n,
               dram_read_bytes=sum(a["rd"] for a in g), dram_write_bytes=sum(a["wr"] for a in g),
               time_share=sum(a["us"] for a in g) / tot, note="ncu --metrics dram__bytes_{read,write}.sum over one "
               "UNet forward at 512x512 (cold cache per launch, so an upper bound on in-graph traffic)")
    print(json.dumps(out))
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], "w"), indent=1)
