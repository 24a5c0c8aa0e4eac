# trip H: fused BN-backward sums -- kernel tests with short timeouts, distill tests, A/B bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "fused_bn_backward or full_size or conv_dgrad or conv_fused" > gpurun_out/gt_h1.log 2>&1; echo "kern -> $?"; tail -5 gpurun_out/gt_h1.log
timeout 600 python -m pytest tests/test_gpu_distill.py -m gpu -q -x --timeout 200 > gpurun_out/gt_h2.log 2>&1; echo "distill -> $?"; tail -5 gpurun_out/gt_h2.log
for f in 0 1; do
GHND_FUSE_BN_REDUCE=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-encode > gpurun_out/bench_f$f.log 2>gpurun_out/bench_f$f.err; echo "bench fuse=$f -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_f$f.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss'])"
done
