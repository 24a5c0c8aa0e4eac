B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-encode --no-config4 --no-stock-torch"
pr() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2: %.1f img/s  e2e %.1f  clocks %s' % (d['value'], d['e2e']['value'], d['clocks']['sm_mhz']))"; }
for rep in 1 2; do
  $B > gpurun_out/ab_new_$rep.log 2>/dev/null; pr gpurun_out/ab_new_$rep.log "late-wait      $rep"
  GHND_S2_SERIAL=1 GHND_S2_SIDE=1 $B > gpurun_out/ab_old_$rep.log 2>/dev/null; pr gpurun_out/ab_old_$rep.log "serial+side    $rep"
  GHND_S2_SIDE=1 $B > gpurun_out/ab_both_$rep.log 2>/dev/null; pr gpurun_out/ab_both_$rep.log "late-wait+side $rep"
done
