# trip J: stem halo mode -- stem tests (hard timeouts), distill tests, A/B bench incl. encode
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout -k 10 240 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 100 -k "stem" > gpurun_out/gt_j1.log 2>&1; echo "kern -> $?"; tail -6 gpurun_out/gt_j1.log
timeout -k 10 400 python -m pytest tests/test_gpu_distill.py -m gpu -q -x --timeout 150 > gpurun_out/gt_j2.log 2>&1; echo "distill -> $?"; tail -3 gpurun_out/gt_j2.log
GHND_NO_STEM_HALO=1 timeout -k 10 500 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h0.log 2>gpurun_out/bench_h0.err; echo "bench halo=0 -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_h0.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss'],d['roofline']['entry_point_ms_per_step']['ghnd_stem_conv_plan_run'],d['encode']['by_batch'])"
timeout -k 10 500 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_h1.log 2>gpurun_out/bench_h1.err; echo "bench halo=1 -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_h1.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['loss'],d['roofline']['entry_point_ms_per_step']['ghnd_stem_conv_plan_run'],d['encode']['by_batch'])"
