# distill tests + short bench (no cpu baseline / encode)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distill.py -m gpu -q -x --timeout 300 > gpurun_out/gt_distill.log 2>&1; echo "distill -> $?"; tail -3 gpurun_out/gt_distill.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-encode > gpurun_out/bench_q.log 2>gpurun_out/bench_q.err; echo "bench -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_q.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'])"
