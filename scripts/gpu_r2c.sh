cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_legacy_abi.py -m gpu -q -x -k "prologue" > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
timeout 900 python scripts/probe_longform_tokens.py > gpurun_out/r2c_tokens.log 2>&1; tail -30 gpurun_out/r2c_tokens.log
PARAKEET_B200_GRAPH=0 timeout 300 python scripts/probe_1stream.py > gpurun_out/r2c_p1.log 2>&1
PARAKEET_B200_GRAPH=0 CHUNKS=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2c_launches_1stream.csv python scripts/probe_1stream.py > gpurun_out/r2c_ncu1.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2c_launches_1stream.csv > gpurun_out/r2c_launch_summary_1stream.csv 2>&1; head -30 gpurun_out/r2c_launch_summary_1stream.csv
PARAKEET_B200_GRAPH=0 STREAMS=128 CHUNKS=5 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2c_launches_128.csv python scripts/probe_1stream.py > gpurun_out/r2c_ncu128.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2c_launches_128.csv > gpurun_out/r2c_launch_summary_128.csv 2>&1; head -30 gpurun_out/r2c_launch_summary_128.csv
