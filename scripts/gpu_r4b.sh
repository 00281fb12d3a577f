# round 2, session 3: ncu --set full of the pre-encode kernels (subsampling stage 1 / stage 2 and the pointwise-conv GEMM) at 1024 streams
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3 --longform 0"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"subsample_stage" -s 180 -c 2 -o gpurun_out/r02g_subsample -f python bench.py $BA > gpurun_out/r4b_ncu1.log 2>&1; echo "ncu subsample rc=$?"
python scripts/ncu_summary.py full gpurun_out/r02g_subsample.ncu-rep > gpurun_out/r02g_subsample_ncu_full_summary.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes" gpurun_out/r02g_subsample_ncu_full_summary.txt | head -12
ls -la gpurun_out/r02g_subsample.ncu-rep
