# long-run stability: 3 x 400 graph-replayed steps on one GPU (catches rare synchronisation faults)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 400 python bench.py --steps 400 --warmup 5 --no-encode --no-cpu-baseline --no-config4 --no-stock-torch > gpurun_out/stress_$i.log 2>gpurun_out/stress_$i.err; echo "stress $i -> $?"; tail -c 120 gpurun_out/stress_$i.log; grep -E "Error" gpurun_out/stress_$i.err | tail -2 | cut -c1-300
done
