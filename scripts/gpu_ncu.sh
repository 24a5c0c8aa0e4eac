# ncu pass only: launch list of one eager step (+ per-layer join), --set full of the top conv shapes
# (report kept, ~20 MB) and of the memory-bound kernels (raw CSV only: gpurun_out/ must stay < 64 MiB).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu list -> $?"
python scripts/join_launches.py gpurun_out/launches.csv gpurun_out/step_ops.json > gpurun_out/per_layer.txt 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launch_summary.txt 2>&1
tail -45 gpurun_out/per_layer.txt
GHND_PROFILE_TOP=${TOPK:-10} timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/conv_top python scripts/profile_step.py > gpurun_out/ncu_top.log 2>&1; echo "ncu full conv -> $?"
ncu -i gpurun_out/conv_top.ncu-rep --page raw --csv > gpurun_out/conv_top_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'sse_kernel|bn_apply_fast|bn_bwd|narrow|maxpool|stem_pack_kernel|wgrad_tc' -f -o /tmp/hbm_kernels python scripts/profile_step.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu full hbm -> $?"
ncu -i /tmp/hbm_kernels.ncu-rep --page raw --csv > gpurun_out/hbm_kernels_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
