# round 2, session 3: transposed 3x3 filter copies for the subsampling kernels -- parity + durations
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_saturated_parity.py tests/test_gpu_model.py tests/test_gpu_offline_long.py -m gpu -x -q > gpurun_out/r4c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r4c_pytest.log
BA="--no-cpu-baseline --no-latency --no-config3 --longform 0"
timeout 600 python bench.py $BA > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err; rc=$?
python - <<PY
import json
d = json.loads(open('gpurun_out/r4c_bench.json').read().strip().splitlines()[-1])
print("bench rc=$rc", round(d['value']), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['achieved']), 'attn', round(d['roofline_hbm'][0]['achieved']), d['clocks']['sm_mhz'])
PY
PARAKEET_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"subsample|gemm_tc_kernel<256, 2>|logmel|audio_append|build_rows" -s 300 -c 24 --csv --log-file gpurun_out/r4c_pre.csv python bench.py --steps 2 --warmup 3 $BA > gpurun_out/r4c_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r4c_pre.csv
