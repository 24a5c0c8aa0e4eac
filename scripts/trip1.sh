cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/t1_smi.txt 2>&1
for k in "quantizer or sse or layout or adam" "bn_ or narrow" "conv_fwd" "conv_dgrad or mixed" "wgrad" "stem"; do
  tag=$(echo "$k" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "$k" > gpurun_out/t1_$tag.log 2>&1
  echo "== $k -> $?" >> gpurun_out/t1_summary.txt
  tail -5 gpurun_out/t1_$tag.log >> gpurun_out/t1_summary.txt
done
cat gpurun_out/t1_summary.txt
