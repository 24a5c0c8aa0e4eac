# ncu --set full (with SASS/source) of selected HBM-bound kernels run by scripts/bench_kernels.py
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python scripts/bench_kernels.py ${BSEL} 2>&1 | tee gpurun_out/bench_kernels.txt
GHND_BENCH_EAGER=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"${KREGEX:-wgrad_narrow_bs|narrow_in_mma|narrow_out_mma|maxpool_bwd|bn_bwd_reduce_fast}" -c ${NCU_COUNT:-12} -f -o gpurun_out/kern python scripts/bench_kernels.py ${BSEL} > gpurun_out/ncu_kern.log 2>&1; echo "ncu -> $?"; tail -3 gpurun_out/ncu_kern.log
ls -la gpurun_out/kern.ncu-rep
