# round 2: tcgen05 attention (split softmax) tests + timings, bench, 1-stream launch list, ncu --set full of the new kernels
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x -s > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -6 gpurun_out/r2e_pytest.log
timeout 300 python scripts/lf_probe.py 3600 2 > gpurun_out/r2e_lf_probe.log 2>&1; tail -2 gpurun_out/r2e_lf_probe.log
PARAKEET_B200_LF_ATTN=1 timeout 300 python scripts/lf_probe.py 3600 2 > gpurun_out/r2e_lf_probe_tma.log 2>&1; tail -1 gpurun_out/r2e_lf_probe_tma.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2e_bench.err
PARAKEET_B200_GRAPH=0 CHUNKS=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2e_launches_1stream.csv python scripts/probe_1stream.py > gpurun_out/r2e_ncu1.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2e_launches_1stream.csv > gpurun_out/r2e_launch_summary_1stream.csv 2>&1; head -12 gpurun_out/r2e_launch_summary_1stream.csv
# full-metric captures: tcgen05 attention at T = 7500 (one launch), then the streaming step's GEMMs at 1024 streams
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lf_attention_tc -c 1 -o gpurun_out/r02_lfattn_tc -f python scripts/lf_probe.py 600 2 > gpurun_out/r2e_ncu2.log 2>&1; echo "ncu lf rc=$?"
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 19200 -c 10 -o gpurun_out/r02_gemm -f python bench.py $BA > gpurun_out/r2e_ncu3.log 2>&1; echo "ncu gemm rc=$?"
for f in r02_lfattn_tc r02_gemm; do python scripts/ncu_summary.py full gpurun_out/$f.ncu-rep > gpurun_out/${f}_ncu_full_summary.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
