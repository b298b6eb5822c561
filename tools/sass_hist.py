"""Opcode histogram of one kernel from an ncu report's SASS source page.
  python tools/sass_hist.py <report.ncu-rep> <kernel regex> [launch-skip]"""
import csv
import subprocess
import sys
from collections import Counter

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
c, s = Counter(), Counter()
tot = 0
for r in data:
    try:
        n = int(r[ie])
    except ValueError:
        continue
    parts = r[isrc].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    c[op] += n
    tot += n
    try:
        s[op] += int(r[iss])
    except ValueError:
        pass
stot = sum(s.values()) or 1
print("total warp instructions", tot, "sass lines", len(data))
for op, n in c.most_common(28):
    print("%-12s %12d %6.2f%%   stall samples %6.2f%%" % (op, n, 100.0 * n / tot, 100.0 * s[op] / stot))
