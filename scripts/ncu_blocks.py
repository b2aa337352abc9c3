"""Coarse stall map of one kernel from an ncu report: the SASS listing cut into blocks of N instructions, with each block's
stall samples, top stall reasons, execution count and notable opcodes (tells which warp role / loop the time goes to).
usage: python scripts/ncu_blocks.py <rep> <kernel-regex> [launch-skip] [block=96] [min-samples=12]"""
import collections
import csv
import re
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
B = int(sys.argv[4]) if len(sys.argv) > 4 else 96
MINS = int(sys.argv[5]) if len(sys.argv) > 5 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print(rows[0][:2], "instructions", len(data), "samples", tot)
NOTE = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "UTCBAR", "BAR", "ATOM", "ATOMG", "RED", "MUFU", "HADD2",
        "F2FP", "STS", "LDS", "LDG", "STG", "FFMA", "SHFL", "DADD", "DMUL", "EXIT")
for b in range(0, len(data), B):
    blk = data[b:b + B]
    s = sum(int(r[iS]) for r in blk)
    if s < MINS:
        continue
    ex = [int(r[iE]) for r in blk if r[iE].isdigit()]
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", r[iSrc].strip()).split()[0].split(".")[0] for r in blk)
    st = collections.Counter()
    for r in blk:
        for i in stall_cols:
            if r[i].isdigit():
                st[hdr[i][6:]] += int(r[i])
    print("%5d-%5d samples %5d (%4.1f%%) exec max %8d  %s | %s" % (b, b + B, s, 100 * s / tot, max(ex) if ex else 0,
                                                                  dict(st.most_common(3)), {k: ops[k] for k in NOTE if k in ops}))
