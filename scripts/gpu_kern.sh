# kernel micro-benchmarks (+ the matching unit tests) -- quick perf iteration trip
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 200 -k "${KSEL:-stem or bn_ or narrow}" > gpurun_out/gt_kern.log 2>&1; echo "tests -> $?"; tail -3 gpurun_out/gt_kern.log
timeout 600 python scripts/bench_kernels.py ${BSEL} 2>&1 | tee gpurun_out/bench_kernels.txt
