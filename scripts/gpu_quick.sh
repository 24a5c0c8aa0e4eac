# focused GPU pass: selected kernel tests, distill tests, bench, launch list of one eager step
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 600 python -m pytest "$@" -m gpu -q -x --timeout 200 > gpurun_out/gt_$tag.log 2>&1
  echo "== $tag -> $?"; tail -4 gpurun_out/gt_$tag.log
}
run kern tests/test_gpu_kernels.py -k "${KSEL:-conv_ or stem or wgrad or bn_}"
run distill tests/test_gpu_distill.py
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench1.log 2>&1; echo "bench -> $?"; tail -1 gpurun_out/bench1.log | cut -c1-400
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu -> $?"
