# Round pass: GPU tests, smoke, bench (with cpu baseline + encode), ncu launch list of one eager step,
# ncu --set full of the top conv shapes and of the memory-bound kernels.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 40 --warmup 5 > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench -> $?"; tail -c 6000 gpurun_out/bench1.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu list -> $?"
GHND_PROFILE_TOP=${TOPK:-10} timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/conv_top python scripts/profile_step.py > gpurun_out/ncu_top.log 2>&1; echo "ncu full conv -> $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'sse_kernel|bn_apply_fast|bn_bwd|narrow|maxpool|stem_pack_kernel' -f -o gpurun_out/hbm_kernels python scripts/profile_step.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu full hbm -> $?"
ls -la gpurun_out
