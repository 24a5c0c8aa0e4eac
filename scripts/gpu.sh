#!/bin/bash
# One parameterised GPU trip (replaces the round-1 gpu_trip_*.sh one-shots).  Run under gpurun:
#   gpurun --timeout 900 -- 'bash scripts/gpu.sh tests "tests/test_gpu_fullsize.py -s" bench "--steps 20 --warmup 5"'
# Verbs (each followed by ONE quoted argument string):
#   tests  <pytest args>      -> gpurun_out/tests_<n>.log      (python -m pytest -m gpu -q <args>)
#   bench  <bench.py args>    -> gpurun_out/bench_<n>.log/.err
#   py     <script + args>    -> gpurun_out/py_<n>.log
#   launches <bench.py args>  -> gpurun_out/launches_<n>.csv   (ncu gpu__time_duration, never a bench value)
#   ab <bench.py args>        -> same-box A/B: hnd_ghnd_object_detectors_b200/libghnd_b200_base.so vs the current build
#   perlayer <batch>          -> gpurun_out/per_layer_<n>.txt + launch_summary_<n>.txt (one eager step under ncu)
#   traffic  <batch>          -> gpurun_out/conv_dram_traffic_<n>.json (dram bytes of the conv_tc launches; copy to
#                                profiles/r2_conv_dram_traffic.json -- bench.py's roofline.traffic reads it)
#   ncu    "<kernel regex>|<skip>|<count>|<cmd>" -> gpurun_out/prof_<n>.ncu-rep (ncu --set full)
mkdir -p gpurun_out
n=0
while [ $# -ge 2 ]; do
  verb=$1; arg=$2; shift 2; n=$((n+1))
  case $verb in
    tests) eval "timeout 1500 python -m pytest -m gpu -q -x $arg" > gpurun_out/tests_$n.log 2>&1; echo "tests_$n rc=$?"; tail -5 gpurun_out/tests_$n.log ;;
    perlayer) # ncu launch list of ONE eager step joined with the plan runs -> per-layer table (arg = batch)
         timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$n.csv python scripts/profile_step.py $arg > gpurun_out/perlayer_$n.log 2>&1; echo "perlayer_$n rc=$?"
         python scripts/join_launches.py gpurun_out/launches_$n.csv gpurun_out/step_ops.json > gpurun_out/per_layer_$n.txt 2>&1
         python scripts/summarize_launches.py gpurun_out/launches_$n.csv 40 > gpurun_out/launch_summary_$n.txt 2>&1; tail -1 gpurun_out/launch_summary_$n.txt ;;
    traffic) # DRAM bytes of every conv_tc_kernel launch of one eager step -> roofline.traffic (arg = batch)
         timeout 1200 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:conv_tc|stem_pool" --csv --log-file gpurun_out/conv_traffic_$n.csv python scripts/profile_step.py $arg > gpurun_out/traffic_$n.log 2>&1; echo "traffic_$n rc=$?"
         python scripts/conv_traffic.py gpurun_out/conv_traffic_$n.csv gpurun_out/conv_dram_traffic_$n.json; cat gpurun_out/conv_dram_traffic_$n.json ;;
    bench) timeout 900 python bench.py $arg > gpurun_out/bench_$n.log 2> gpurun_out/bench_$n.err; echo "bench_$n rc=$?"; head -c 600 gpurun_out/bench_$n.log; echo ;;
    py) timeout 900 python $arg > gpurun_out/py_$n.log 2>&1; echo "py_$n rc=$?"; tail -30 gpurun_out/py_$n.log ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$n.csv python bench.py $arg > gpurun_out/launches_$n.log 2>&1; echo "launches_$n rc=$?" ;;
    ncu) IFS='|' read -r rx skip cnt cmd <<< "$arg"
         timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/prof_$n $cmd > gpurun_out/ncu_$n.log 2>&1; echo "ncu_$n rc=$?" ;;
    ab) # same-box A/B of two builds of the library: libghnd_b200_base.so (A) vs libghnd_b200.so (B), alternating
         for rep in 1 2; do
           for v in base new; do
             if [ $v = base ]; then export GHND_LIB_PATH=$PWD/hnd_ghnd_object_detectors_b200/libghnd_b200_base.so; else unset GHND_LIB_PATH; fi
             timeout 600 python bench.py --no-cpu-baseline --no-encode --no-config4 $arg > gpurun_out/ab_${n}_${v}_$rep.log 2> gpurun_out/ab_${n}_${v}_$rep.err
             python -c "import json,sys; d=json.loads(open('gpurun_out/ab_${n}_${v}_$rep.log').read().strip().splitlines()[-1]); print('$v $rep: %.1f img/s  e2e %.1f  conv frac %.3f  clocks %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz']))"
           done
         done; unset GHND_LIB_PATH ;;
    *) echo "unknown verb $verb" ;;
  esac
done
