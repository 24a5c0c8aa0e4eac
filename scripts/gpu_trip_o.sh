# trip O: packed-half2 epilogue -- tests, microbench A/B, bench A/B
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 100 -k "conv_ or stem or dgrad" > gpurun_out/gt_o1.log 2>&1; echo "kern -> $?"; tail -6 gpurun_out/gt_o1.log
timeout -k 10 400 python -m pytest tests/test_gpu_distill.py -m gpu -q --timeout 150 > gpurun_out/gt_o2.log 2>&1; echo "distill -> $?"; tail -4 gpurun_out/gt_o2.log
for b in ${HALF_LEVELS:-0 1 2}; do
echo "== GHND_EPI_HALF=$b"; GHND_EPI_HALF=$b timeout 120 python scripts/bench_kernels.py stem 2>&1 | tail -5
GHND_EPI_HALF=$b timeout -k 10 500 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_x$b.log 2>gpurun_out/bench_x$b.err; echo "bench half=$b -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_x$b.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss'],d['roofline']['entry_point_ms_per_step']['ghnd_stem_conv_plan_run'],d['roofline']['entry_point_ms_per_step']['ghnd_conv_plan_run'],d['encode']['by_batch'])"
done
