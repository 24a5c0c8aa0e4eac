#!/bin/bash
# same-box A/B of an environment switch: ab_env.sh "VAR=a" "VAR=b" [reps]   (bench.py, device-resident step)
B="python bench.py --steps ${STEPS:-40} --warmup 5 --no-cpu-baseline --no-encode --no-config4 --no-stock-torch"
mkdir -p gpurun_out
for rep in $(seq 1 ${3:-2}); do
  for v in "$1" "$2"; do
    env $v $B > gpurun_out/abenv.log 2>/dev/null
    python -c "import json; d=json.loads(open('gpurun_out/abenv.log').read().strip().splitlines()[-1]); print('%-28s %d: %.1f img/s  e2e %.1f  conv %.3f ms  frac %.3f  clocks %s' % ('$v', $rep, d['value'], d['e2e']['value'], d['roofline']['entry_point_ms_per_step'].get('ghnd_conv_plan_run', 0), d['roofline']['frac'], d['clocks']['sm_mhz']))"
  done
done
