# trip F: full GPU tests + bench (default args incl. cpu baseline + encode)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench1.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']);print(json.dumps(d['roofline_hbm_kernels'],indent=1));print(d['encode']);print(d['cpu_baseline']);print(d['roofline']['entry_point_ms_per_step'])"
tail -3 gpurun_out/bench1.err
