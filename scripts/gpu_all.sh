cd ${GRAFT_REPO_ROOT:-.}
bash scripts/gpu_tests.sh
timeout 1200 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_last.log 2>&1; echo "bench -> $?"; tail -1 gpurun_out/bench_last.log | python -c "
import sys, json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'])
except Exception as e: print(l[-2000:])
"
