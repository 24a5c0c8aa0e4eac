"""Sum dram__bytes_read/write over the conv_tc_kernel (+ stem_pool_kernel) launches of one eager GHND step (ncu CSV written by
`scripts/gpu.sh traffic <batch>`) -> the JSON that bench.py reports as roofline.traffic.
Usage: conv_traffic.py conv_traffic.csv out.json"""
import csv
import json
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows))
hdr, r = r[0], r[1:]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "nsecond": 1e-9,
         "msecond": 1e-3, "ms": 1e-3}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
ids = set()
for x in r:
    if not ("conv_tc_kernel" in x[ki] or "stem_pool_kernel" in x[ki]) or x[mi] not in tot:
        continue
    ids.add(x[ii])
    tot[x[mi]] += float(x[vi].replace(",", "")) * scale.get(x[ui], 1.0)
out = {"conv_tc_dram_bytes_per_step": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
       "read": tot["dram__bytes_read.sum"], "write": tot["dram__bytes_write.sum"], "launches": len(ids),
       "kernel_time_s": tot["gpu__time_duration.sum"],
       "note": "sum over the %d conv_tc_kernel / stem_pool_kernel launches of one eager GHND step (batch 4) from ncu --metrics "
               "dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (scripts/gpu.sh traffic 4); per "
               "step, like achieved" % len(ids)}
json.dump(out, open(sys.argv[2], "w"), indent=1)
