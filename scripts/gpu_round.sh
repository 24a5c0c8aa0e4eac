# Full GPU pass: tests (separate processes), smoke, bench, ncu launch list of one eager step.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1.log 2>&1; echo "bench -> $?"; tail -1 gpurun_out/bench1.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu -> $?"
