"""Per-shape roofline bounds of the tensor-core conv launches of one step.
Usage: conv_bounds.py per_layer.txt   (the aggregated table written by join_launches.py)
For each shape: HBM time of the compulsory bytes (src + dst + residual/mask/acc operands, 16-bit) at the
measured copy bandwidth, tensor time at the measured sustained bf16 peak, the larger = bound, and the
slack = measured - bound."""
import json, os, re, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(root, "MEASURED_PEAKS.json")) else {}
HBM = pk.get("hbm_gbs", 6650.0) * 1e9
TC = pk.get("bf16_tflops_sustained", 1400.0) * 1e12
rows = []
seen = False
for line in open(sys.argv[1]):
    if "aggregated" in line:
        seen = True
        continue
    if not seen:
        continue
    m = re.match(r"(fwd|dgrad|wgrad) N(\d+) (\d+)x(\d+) C(\d+) K(\d+) (\d)x(\d) (?:s(\d) )?p(\d)(.*?)\s+x(\d+)\s+([\d.]+) us", line)
    if not m:
        continue
    kind, N, H, W, C, K, R, S, st, pad, flags, cnt, us = m.groups()
    N, H, W, C, K, R, S, pad, cnt = map(int, (N, H, W, C, K, R, S, pad, cnt))
    st = int(st or 1)
    us = float(us) / cnt
    Ho, Wo = (H + 2 * pad - R) // st + 1, (W + 2 * pad - S) // st + 1
    flops = 2.0 * N * Ho * Wo * K * C * R * S
    if kind == "fwd":
        src, dst = N * H * W * C, N * Ho * Wo * K
    elif kind == "dgrad":
        src, dst = N * Ho * Wo * K, N * H * W * C
    else:
        src, dst = N * H * W * C + N * Ho * Wo * K, 0
    n_in = ("+res" in flags) + ("mask" in flags) + ("acc" in flags)
    byts = 2.0 * (src + dst * (1 + n_in))
    t_h, t_t = byts / HBM * 1e6, flops / TC * 1e6
    rows.append((us * cnt, cnt, line[:52].strip(), us, t_h, t_t, max(t_h, t_t)))
tot = sum(r[0] for r in rows)
slack = sum((r[3] - r[6]) * r[1] for r in rows)
print("%-52s %3s %8s %8s %8s %6s %8s" % ("shape", "n", "us/launch", "hbm us", "tensor us", "bound%", "slack us"))
for t, cnt, name, us, th, tt, b in sorted(rows, key=lambda r: -(r[3] - r[6]) * r[1]):
    print("%-52s %3d %8.1f %8.1f %8.1f %5.0f%% %8.1f" % (name, cnt, us, th, tt, 100 * b / us, (us - b) * cnt))
print("total %.0f us, sum of bounds %.0f us, slack %.0f us" % (tot, tot - slack, slack))
