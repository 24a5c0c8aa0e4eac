"""Top stalled SASS instructions of an `ncu --page source --csv` dump.  Usage: ncu_hot.py src.csv [n]"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = r[1]; rows = r[2:]
si = hdr.index("# Samples"); src = hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(x[si] or 0) for x in rows)
print("total samples", tot)
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][si] or 0))[:n]
for i in sorted(order):
    x = rows[i]
    st = sorted(((int(x[j] or 0), hdr[j][6:]) for j in stall), reverse=True)[:3]
    print("%5d %5.1f%% line%5d  %-70s %s" % (int(x[si]), 100.0 * int(x[si]) / tot, i, x[src].strip()[:70],
                                       " ".join("%s:%d" % (b, a) for a, b in st if a)))
