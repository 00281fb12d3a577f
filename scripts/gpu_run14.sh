cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -15
  echo "gemm rc=$?"
  timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
  BA="--steps 5 --warmup 3 --no-cpu-baseline --no-latency"
  for bn in 0 128 256; do
    PARAKEET_B200_GEMM_BN=$bn timeout 600 python bench.py $BA > gpurun_out/b14_$bn.json 2> gpurun_out/b14_$bn.err; echo "bench bn=$bn rc=$?"
    python -c "import json;d=json.load(open('gpurun_out/b14_$bn.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'],d['roofline']['frac'])"
  done
) > gpurun_out/run14.log 2>&1
tail -40 gpurun_out/run14.log
