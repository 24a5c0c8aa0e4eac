# 2-GPU data-parallel bench (NCCL all-reduce of the flat student gradient) + the reference arm launched the same way.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N -> $?"
tail -c 3000 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
