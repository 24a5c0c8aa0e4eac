# trip E: full GPU tests + kernel microbench + bench with and without the side stream
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 300 python scripts/bench_kernels.py maxpool narrow 2>&1 | tee gpurun_out/bench_kernels.txt
GHND_SIDE_STREAM=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-encode > gpurun_out/bench_noside.log 2>gpurun_out/bench_noside.err; echo "bench noside -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_noside.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-encode > gpurun_out/bench_side.log 2>gpurun_out/bench_side.err; echo "bench side -> $?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_side.log').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'])"
tail -3 gpurun_out/bench_side.err
