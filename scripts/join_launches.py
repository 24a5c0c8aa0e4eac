"""Join an ncu launch list (gpu__time_duration.sum CSV of scripts/profile_step.py) with
gpurun_out/step_ops.json -> per-layer table: time, TFLOP/s.  Usage: join_launches.py launches.csv step_ops.json"""
import csv, json, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows)); hdr, r = r[0], r[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
import re


def kname(full):
    """'void conv_tc_kernel<1, 0, 0>(ghnd::ConvKernelParams)' -> 'conv_tc_kernel' (the template
    instantiations of one kernel share a launch queue)."""
    n = full.split("(")[0]
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"<.*$", "", n)
    return n.split("::")[-1]


launches = [(kname(x[ki]), float(x[vi].replace(",", "")) / 1e3) for x in r]
ops = json.load(open(sys.argv[2]))
queues = {}
for name, us in launches:
    queues.setdefault(name, []).append(us)
pos = {k: 0 for k in queues}
tot = sum(us for _, us in launches)
agg = {}
print("%-58s %3s %9s %8s" % ("plan run", "n", "us", "TFLOP/s"))
for o in ops:
    k = o["kernel"]; n = o["launches"]
    us = sum(queues[k][pos[k]:pos[k] + n]); pos[k] += n
    print("%-58s %3d %9.1f %8.1f" % (o["desc"], n, us, o["flops"] / us / 1e6))
    a = agg.setdefault(o["desc"], [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += o["flops"]
print("\n== aggregated by shape, sorted by time (total of all kernels %.1f us) ==" % tot)
for d, (c, us, fl) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-58s x%-3d %9.1f us %5.1f%% %8.1f TFLOP/s" % (d, c, us, 100 * us / tot, fl / us / 1e6))
