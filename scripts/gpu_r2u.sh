# round 2: calibrate the tail-slice cost (kbench), dwconv4 rewrite check, bench A/B
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
for t in 1 2 4; do
  echo "== tail max $t (T2=T4=1: any cut that fits)"
  for shape in "6144 4096 1024 50 silu" "6144 1024 1024 50 partial1pb" "6144 1024 4096 50 partial2pb" "6144 1024 4096 50 partial1pb" "6144 2048 1024 50 silu" "6144 3072 1024 50 silu"; do
    PARAKEET_B200_GEMM_BN=512 PARAKEET_B200_GEMM_TAIL=$t PARAKEET_B200_GEMM_TAIL_T2=1 PARAKEET_B200_GEMM_TAIL_T4=1 timeout 60 $K gemm $shape 0 8 2>&1 | tail -1
  done
done > gpurun_out/r2u_kbench.txt 2>&1; cat gpurun_out/r2u_kbench.txt
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_set.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log; tail -3 gpurun_out/r2u_pytest.log
BB="--no-cpu-baseline --no-config3 --longform 0 --no-latency"
for t in 4 2 0; do PARAKEET_B200_GEMM_TAIL=$t timeout 600 python bench.py $BB > gpurun_out/r2u_bench_tail$t.json 2> gpurun_out/r2u_bench_tail$t.err; done
PARAKEET_B200_DWCONV4=0 PARAKEET_B200_GEMM_TAIL=0 timeout 600 python bench.py $BB > gpurun_out/r2u_bench_tail0_dwold.json 2> gpurun_out/r2u_bench_tail0_dwold.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2u_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],3), 'rtfx', round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
