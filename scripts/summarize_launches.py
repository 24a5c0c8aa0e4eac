"""Per-kernel totals of an ncu gpu__time_duration launch list.  Usage: summarize_launches.py launches.csv [n]"""
import csv, collections, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
r = list(csv.reader(rows)); hdr, r = r[0], r[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict(); tot = 0
for x in r:
    k = x[ki].split("(")[0]; v = float(x[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
for k, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1])[:n]:
    print("%-44s %4d %9.1f us %5.1f%%" % (k, c, v / 1e3, 100 * v / tot))
print("total %.1f us over %d launches" % (tot / 1e3, len(r)))
