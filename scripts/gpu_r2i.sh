cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for cfg in "1 0" "1 20000" "2 0" "2 20000"; do set -- $cfg; PARAKEET_B200_LF_SPLIT=$1 PARAKEET_B200_LF_HINT_NS=$2 timeout 300 python scripts/lf_probe.py 3600 2 > gpurun_out/r2i_lf_probe_s$1_h$2.log 2>&1; echo "split $1 hint $2: $(tail -1 gpurun_out/r2i_lf_probe_s$1_h$2.log)"; done
timeout 900 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_pytest.log
PARAKEET_B200_LF_HINT_NS=20000 timeout 600 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x -k "ragged or tiny or tcgen05" > gpurun_out/r2i_pytest_hint.log 2>&1; tail -2 gpurun_out/r2i_pytest_hint.log
BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --no-config3"
PARAKEET_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 36000 -c 900 --csv --log-file gpurun_out/r2i_launches_1024.csv python bench.py $BA > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2i_launches_1024.csv > gpurun_out/r2i_launch_summary_1024.csv 2>&1; head -16 gpurun_out/r2i_launch_summary_1024.csv
