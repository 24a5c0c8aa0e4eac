# Round-1 trip D: new transform/multi-scale tests + the full GPU suite, smoke, short bench.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-encode > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench -> $?"; tail -c 1500 gpurun_out/bench1.log; tail -5 gpurun_out/bench1.err
