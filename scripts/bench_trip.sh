cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distill.py -m gpu -q -x --timeout 300 -s > gpurun_out/gt_distill.log 2>&1
echo "distill -> $?"; tail -4 gpurun_out/gt_distill.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.log 2>&1; echo "bench -> $?"; tail -3 gpurun_out/bench1.log
