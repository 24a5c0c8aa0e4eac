# trip N: epilogue TMEM-load prefetch -- tests (hard timeouts), A/B bench + per-layer lists
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 100 -k "conv_ or stem or dgrad" > gpurun_out/gt_n1.log 2>&1; echo "kern -> $?"; tail -4 gpurun_out/gt_n1.log
timeout -k 10 400 python -m pytest tests/test_gpu_distill.py -m gpu -q -x --timeout 150 > gpurun_out/gt_n2.log 2>&1; echo "distill -> $?"; tail -3 gpurun_out/gt_n2.log
for b in 0 1; do
GHND_EPI_PREFETCH=$b timeout -k 10 500 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_p$b.log 2>gpurun_out/bench_p$b.err; echo "bench prefetch=$b -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_p$b.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss'],d['roofline']['entry_point_ms_per_step']['ghnd_stem_conv_plan_run'],d['roofline']['entry_point_ms_per_step']['ghnd_conv_plan_run'],d['encode']['by_batch'])"
GHND_EPI_PREFETCH=$b timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_p$b.csv python scripts/profile_step.py > gpurun_out/ncu_step_p$b.log 2>&1
python scripts/join_launches.py gpurun_out/launches_p$b.csv gpurun_out/step_ops.json > gpurun_out/per_layer_p$b.txt 2>&1
done
