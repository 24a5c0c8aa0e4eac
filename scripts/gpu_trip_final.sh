# Round-1 final evidence (after the packed epilogues / templated conv kernels)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke -> $?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 40 --warmup 5 --encode-sweep > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench -> $?"; tail -c 400 gpurun_out/bench1.log; tail -2 gpurun_out/bench1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "bench ref -> $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu list -> $?"
python scripts/join_launches.py gpurun_out/launches.csv gpurun_out/step_ops.json > gpurun_out/per_layer.txt 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv 40 > gpurun_out/launch_summary.txt 2>&1
head -12 gpurun_out/launch_summary.txt; tail -1 gpurun_out/launch_summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_encode64.csv python scripts/profile_encode.py 64 > gpurun_out/ncu_enc.log 2>&1; echo "ncu encode list -> $?"
GHND_PROFILE_TOP=10 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'conv_tc' -f -o /tmp/conv_top python scripts/profile_step.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv top10 -> $?"
ncu -i /tmp/conv_top.ncu-rep --page raw --csv > gpurun_out/conv_top10_raw.csv 2>/dev/null
timeout 300 python scripts/bench_kernels.py stem > gpurun_out/bench_kernels_conv.txt 2>&1; cat gpurun_out/bench_kernels_conv.txt
