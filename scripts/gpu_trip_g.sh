# Round-1 trip G (final evidence): bench (default args), encode sweep, launch list + per-layer join,
# ncu --set full of the streaming kernels and of the encode path, reference arm sanity.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 40 --warmup 5 --encode-sweep > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench -> $?"; tail -c 600 gpurun_out/bench1.log; tail -3 gpurun_out/bench1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "bench ref -> $?"; tail -c 1200 gpurun_out/bench_ref.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu list -> $?"
python scripts/join_launches.py gpurun_out/launches.csv gpurun_out/step_ops.json > gpurun_out/per_layer.txt 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launch_summary.txt 2>&1
head -45 gpurun_out/launch_summary.txt
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'narrow|sse_kernel|bn_bwd|bn_apply|maxpool|wgrad_tc|stem_wgrad' -f -o /tmp/hbm_kernels python scripts/profile_step.py > gpurun_out/ncu_hbm.log 2>&1; echo "ncu full hbm -> $?"
ncu -i /tmp/hbm_kernels.ncu-rep --page raw --csv > gpurun_out/hbm_kernels_raw.csv 2>/dev/null
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_encode64.csv python scripts/profile_encode.py 64 > gpurun_out/ncu_enc.log 2>&1; echo "ncu encode list -> $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'quant|narrow' -f -o /tmp/enc_kernels python scripts/profile_encode.py 64 > gpurun_out/ncu_enc_full.log 2>&1; echo "ncu full encode -> $?"
ncu -i /tmp/enc_kernels.ncu-rep --page raw --csv > gpurun_out/encode_kernels_raw.csv 2>/dev/null
GHND_PROFILE_TOP=10 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'conv_tc' -f -o /tmp/conv_top python scripts/profile_step.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv top10 -> $?"
ncu -i /tmp/conv_top.ncu-rep --page raw --csv > gpurun_out/conv_top10_raw.csv 2>/dev/null
du -sh gpurun_out; ls gpurun_out
