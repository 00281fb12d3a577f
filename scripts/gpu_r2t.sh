# round 2: tail column slices in the CTA-pair GEMM: GEMM tests, parity subset, kbench, bench A/B, 128-stream launch list
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/r2t_pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_gemm.log; tail -3 gpurun_out/r2t_pytest_gemm.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py tests/test_gpu_full_size.py -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log; tail -3 gpurun_out/r2t_pytest.log
K=trt-asr-engine_b200/bin/kbench
for t in 1 0; do
  echo "== tail $t"
  PARAKEET_B200_GEMM_TAIL=$t timeout 60 $K gemm 6144 4096 1024 50 silu 0 8 2>&1 | tail -1
  PARAKEET_B200_GEMM_TAIL=$t timeout 60 $K gemm 6144 1024 1024 50 partial1pb 0 8 2>&1 | tail -1
  PARAKEET_B200_GEMM_TAIL=$t timeout 60 $K gemm 6144 1024 4096 50 partial2pb 0 8 2>&1 | tail -1
  PARAKEET_B200_GEMM_TAIL=$t timeout 60 $K gemm 6144 1024 4096 50 partial1pb 0 8 2>&1 | tail -1
done > gpurun_out/r2t_kbench.txt 2>&1; cat gpurun_out/r2t_kbench.txt
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
timeout 600 python bench.py $BB > gpurun_out/r2t_bench_new.json 2> gpurun_out/r2t_bench_new.err
PARAKEET_B200_GEMM_TAIL=0 timeout 600 python bench.py $BB > gpurun_out/r2t_bench_notail.json 2> gpurun_out/r2t_bench_notail.err
timeout 600 python bench.py $BB > gpurun_out/r2t_bench_new2.json 2> gpurun_out/r2t_bench_new2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
PARAKEET_B200_GRAPH=0 STREAMS=128 CHUNKS=92 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 35500 -c 1200 --csv --log-file gpurun_out/r2t_launches_128.csv python scripts/probe_1stream.py > gpurun_out/r2t_ncu128.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2t_launches_128.csv > gpurun_out/r02_launch_summary_128.csv 2>&1; head -30 gpurun_out/r02_launch_summary_128.csv
