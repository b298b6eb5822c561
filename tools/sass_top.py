"""Top SASS instructions of one kernel by stall samples (ncu report, SASS source page).
  python tools/sass_top.py <report.ncu-rep> <kernel regex> [launch-skip] [n]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "1"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iss]) for r in data if r[iss].isdigit()) or 1
idx = sorted(range(len(data)), key=lambda k: -(int(data[k][iss]) if data[k][iss].isdigit() else 0))[:n]
for k in sorted(idx):
    r = data[k]
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, h[6:]) for i, h in stall_cols), reverse=True)[:2]
    print("%5d %5.2f%% ex=%9s  %-70s %s" % (k, 100.0 * int(r[iss]) / tot, r[ie], r[isrc][:70], " ".join("%s:%d" % (h, v) for v, h in st if v)))
