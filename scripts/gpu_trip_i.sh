# trip I: dual stem -- stem tests, distill tests, A/B bench
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "stem or full_size" > gpurun_out/gt_i1.log 2>&1; echo "kern -> $?"; tail -5 gpurun_out/gt_i1.log
timeout 600 python -m pytest tests/test_gpu_distill.py -m gpu -q -x --timeout 200 > gpurun_out/gt_i2.log 2>&1; echo "distill -> $?"; tail -3 gpurun_out/gt_i2.log
for f in 0 1; do
GHND_DUAL_STEM=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-encode > gpurun_out/bench_d$f.log 2>gpurun_out/bench_d$f.err; echo "bench dual=$f -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_d$f.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss'],d['roofline']['entry_point_ms_per_step']['ghnd_stem_conv_plan_run'])"
done
