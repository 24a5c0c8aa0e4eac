# Round-1 final evidence trip: full GPU tests, smoke, bench (default args + encode sweep), reference arm,
# launch list + per-layer join, ncu --set full of the streaming / encode / top-10 conv kernels.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 40 --warmup 5 --encode-sweep > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench -> $?"; tail -c 700 gpurun_out/bench1.log; tail -2 gpurun_out/bench1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "bench ref -> $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu list -> $?"
python scripts/join_launches.py gpurun_out/launches.csv gpurun_out/step_ops.json > gpurun_out/per_layer.txt 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launch_summary.txt 2>&1
head -42 gpurun_out/launch_summary.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'narrow|sse_kernel|bn_bwd|bn_apply|maxpool|wgrad_tc|stem_wgrad' -f -o /tmp/hbm_kernels python scripts/profile_step.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu full hbm -> $?"
ncu -i /tmp/hbm_kernels.ncu-rep --page raw --csv > gpurun_out/hbm_kernels_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_encode64.csv python scripts/profile_encode.py 64 > gpurun_out/ncu_enc.log 2>&1; echo "ncu encode list -> $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'quant|narrow' -f -o /tmp/enc_kernels python scripts/profile_encode.py 64 > gpurun_out/ncu_enc_full.log 2>&1; echo "ncu full encode -> $?"
ncu -i /tmp/enc_kernels.ncu-rep --page raw --csv > gpurun_out/encode_kernels_raw.csv 2>/dev/null
GHND_PROFILE_TOP=10 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'conv_tc' -f -o /tmp/conv_top python scripts/profile_step.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv top10 -> $?"
ncu -i /tmp/conv_top.ncu-rep --page raw --csv > gpurun_out/conv_top10_raw.csv 2>/dev/null
timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc_kernel -c 4 -f -o /tmp/stem python scripts/profile_step.py > gpurun_out/ncu_stem.log 2>&1; echo "ncu stem -> $?"
ncu -i /tmp/stem.ncu-rep --page raw --csv > gpurun_out/stem_raw.csv 2>/dev/null
timeout 300 python scripts/bench_kernels.py > gpurun_out/bench_kernels.txt 2>&1; tail -40 gpurun_out/bench_kernels.txt
du -sh gpurun_out
