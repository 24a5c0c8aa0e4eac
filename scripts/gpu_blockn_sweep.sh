# per-layer launch lists with the tile width forced to 64 / 128 / 256 and with the cost model's choice
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for bn in 0 64 128 256; do
  if [ "$bn" = "0" ]; then unset GHND_BLOCK_N; else export GHND_BLOCK_N=$bn; fi
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bn$bn.csv python scripts/profile_step.py > gpurun_out/ncu_bn$bn.log 2>&1; echo "bn=$bn -> $?"
  python scripts/join_launches.py gpurun_out/launches_bn$bn.csv gpurun_out/step_ops.json > gpurun_out/per_layer_bn$bn.txt 2>&1
done
