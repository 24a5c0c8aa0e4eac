cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for bn in 64 128 256; do
GHND_BLOCK_N=$bn timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bn$bn.csv python scripts/profile_step.py > gpurun_out/ncu_bn$bn.log 2>&1; echo "bn $bn -> $?"
cp gpurun_out/step_ops.json gpurun_out/step_ops_bn$bn.json
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "default -> $?"
