# bench line (default args) + DRAM traffic of every conv_tc launch of one eager step
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err; echo "bench -> $?"; tail -c 3000 gpurun_out/bench_full.log; tail -5 gpurun_out/bench_full.err
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/conv_traffic.csv -k regex:conv_tc python scripts/profile_step.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu -> $?"
