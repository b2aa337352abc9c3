"""Top stall locations of one kernel from `ncu -i rep --page source --csv` output (SASS view).
usage: python scripts/ncu_stalls.py <rep> <kernel-regex> [launch-skip] [topN]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip",
                          skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    print(rows[0][:2])
    i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for r in data if r[i_s].isdigit())
    print("total samples", tot, "instructions", len(data))
    agg = {}
    for r in data:
        for i in stall_cols:
            if r[i].isdigit():
                agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
    print("stall totals:", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(data, key=lambda r: -int(r[i_s]) if r[i_s].isdigit() else 0)[:topn]
    for r in top:
        st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{int(r[i_s]):7d} {100 * int(r[i_s]) / max(tot, 1):5.1f}%  {r[i_src].strip()[:80]:80s} {st}")


if __name__ == "__main__":
    main()
