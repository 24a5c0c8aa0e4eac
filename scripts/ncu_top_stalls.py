"""Top stall-sample SASS lines of one kernel from an .ncu-rep (needs --import-source on / --set full).
    python scripts/ncu_top_stalls.py report.ncu-rep kernel-regex [top-n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > ie:
        body.append(r)
tot = sum(int(r[si]) for r in body)
execd = sum(int(r[ie]) for r in body)
print(rows[0][1][:100])
print("samples %d, SASS lines %d, warp-instructions executed %d" % (tot, len(body), execd))
top = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:topn]
for i in sorted(top):
    print("%5d %-74s %6s %9s" % (i, body[i][1][:74], body[i][si], body[i][ie]))
