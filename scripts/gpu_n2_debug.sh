cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
GHND_NO_GRAPH=1 CUDA_LAUNCH_BLOCKING=1 GHND_BENCH_LR=0 timeout 400 python bench.py --steps 400 --warmup 5 --no-encode --no-cpu-baseline > gpurun_out/n1_e$i.log 2>gpurun_out/n1_e$i.err; echo "E$i -> $?"; grep -E "GhndError:" gpurun_out/n1_e$i.err | tail -1 | cut -c1-400
done
