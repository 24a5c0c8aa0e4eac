# Runs the GPU test groups in separate processes (a faulting kernel poisons its CUDA context).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
rm -f gpurun_out/gt_summary.txt
run() {
  tag=$1; shift
  timeout 900 python -m pytest "$@" -m gpu -q -x --timeout 300 -s > gpurun_out/gt_$tag.log 2>&1
  echo "== $tag -> $?" >> gpurun_out/gt_summary.txt
  tail -4 gpurun_out/gt_$tag.log >> gpurun_out/gt_summary.txt
}
run basic tests/test_gpu_kernels.py -k "quantizer or sse or layout or adam or bn_ or narrow or stem"
run conv tests/test_gpu_kernels.py -k "conv_"
run wgrad tests/test_gpu_kernels.py -k "wgrad"
run distill tests/test_gpu_distill.py
cat gpurun_out/gt_summary.txt
