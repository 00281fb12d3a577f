cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
for sp in 1 2; do PARAKEET_B200_LF_SPLIT=$sp timeout 300 python scripts/lf_probe.py 3600 2 > gpurun_out/r2g_lf_probe_split$sp.log 2>&1; tail -1 gpurun_out/r2g_lf_probe_split$sp.log; done
timeout 1500 python -m pytest tests/test_gpu_offline_long.py tests/test_gpu_model.py tests/test_gpu_parity_set.py tests/test_gpu_full_size.py tests/test_gpu_gemm.py -m gpu -q -x -s > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -6 gpurun_out/r2g_pytest.log
PARAKEET_B200_LF_SPLIT=2 timeout 600 python -m pytest tests/test_gpu_offline_long.py -m gpu -q -x -k "ragged or tiny or tcgen05" > gpurun_out/r2g_pytest_split2.log 2>&1; tail -2 gpurun_out/r2g_pytest_split2.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
for k in ('config3_64streams_mixed_cache','config5_longform','latency_1stream'): print(k, d.get(k))
PY
