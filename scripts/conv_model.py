"""Analytic bound model of the conv launches of one step (from gpurun_out/step_ops.json + launch list):
per launch tensor / smem-ingest / HBM bounds under tiling options, next to the measured time."""
import json, re, sys, math, csv
ops = json.load(open("gpurun_out/step_ops.json"))
rows = [l for l in open("gpurun_out/launches.csv") if l.startswith('"')]
r = list(csv.reader(rows)); hdr, r = r[0], r[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
q = {}
for x in r: q.setdefault(x[ki].split("(")[0], []).append(float(x[vi].replace(",", "")) / 1e3)
pos = {k: 0 for k in q}
CLK = 1.9e9; SMS = 148; ING = 48.0; HBM = 6.4e12
tot = {"meas": 0, "cur": 0, "m256": 0, "halo": 0, "both": 0, "hbm": 0, "tensor": 0}
out = []
for o in ops:
    k = o["kernel"]; n = o["launches"]; us = sum(q[k][pos[k]:pos[k] + n]); pos[k] += n
    m = re.match(r"(fwd|dgrad) N(\d+) (\d+)x(\d+) C(\d+) K(\d+) (\d)x(\d) s(\d) p(\d)(.*)", o["desc"])
    if not m: continue
    kind, N, H, W, C, K, R, S, st, pad, rest = m.groups()
    N, H, W, C, K, R, S, st, pad = map(int, (N, H, W, C, K, R, S, st, pad))
    Ho, Wo = (H + 2 * pad - R) // st + 1, (W + 2 * pad - S) // st + 1
    if kind == "fwd": cin, cout, px = C, K, N * Ho * Wo
    else: cin, cout, px = K, C, N * H * W
    taps = R * S if not (kind == "dgrad" and st == 2) else R * S / 4.0
    mt = math.ceil(px / 128)
    def bound(bn, m256, halo):
        nt = cout // bn
        units = taps * cin / 64
        a = 16384 * (2 if m256 else 1); b = bn * 128
        a_units = units if not halo or taps == 1 else units / taps * 1.45
        tiles = math.ceil(mt / (2 if m256 else 1)) * nt
        waves = math.ceil(tiles / SMS)
        ing = waves * (a_units * a + units * b) / ING / CLK
        ten = waves * units * 4 * (bn / 2) * (2 if m256 else 1) / CLK
        return max(ing, ten) * 1e6
    bns = [b for b in (64, 128, 256) if cout % b == 0]
    cur = min(bound(b, False, False) for b in bns)
    m256 = min(cur, min(bound(b, True, False) for b in bns))
    halo = min(bound(b, False, True) for b in bns)
    both = min(halo, m256, min(bound(b, True, True) for b in bns))
    nin = (1 if "+res" in rest or "acc" in rest else 0) + (1 if "mask" in rest else 0)
    hbm = (px * cout * 2 * (1 + nin) + (N * H * W * C if kind == "fwd" else N * Ho * Wo * K) * 2 / (1 if st == 1 or kind == "dgrad" else 1)) / HBM * 1e6
    ten = o["flops"] / (SMS * 8192 * CLK) * 1e6
    for key, v in (("meas", us), ("cur", max(cur, hbm)), ("m256", max(m256, hbm)), ("halo", max(halo, hbm)), ("both", max(both, hbm)), ("hbm", hbm), ("tensor", ten)):
        tot[key] += v
    out.append((us, o["desc"], cur, m256, halo, both, hbm, ten))
print("totals (us):", {k: round(v) for k, v in tot.items()})
agg = {}
for us, d, cur, m256, halo, both, hbm, ten in out:
    a = agg.setdefault(d, [0, 0, 0, 0, 0, 0, 0, 0]); a[0] += 1
    for i, v in enumerate((us, cur, m256, halo, both, hbm, ten)): a[i + 1] += v
print("%-52s %3s %7s %7s %7s %7s %7s %7s %7s" % ("shape", "n", "meas", "cur", "m256", "halo", "both", "hbm", "tensor"))
for d, a in sorted(agg.items(), key=lambda t: -t[1][1])[:int(sys.argv[1]) if len(sys.argv) > 1 else 30]:
    print("%-52s %3d %7.0f %7.0f %7.0f %7.0f %7.0f %7.0f %7.0f" % (d, *a))
