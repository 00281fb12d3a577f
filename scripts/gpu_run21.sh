cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15
  echo "model rc=$?"
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
  for s in 1024 128; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --streams $s > gpurun_out/b21_$s.json 2> gpurun_out/b21_$s.err; echo "bench rc=$?"; tail -3 gpurun_out/b21_$s.err
  python -c "import json;d=json.load(open('gpurun_out/b21_$s.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'],d['roofline']['frac'])"
  done
) > gpurun_out/run21.log 2>&1
tail -30 gpurun_out/run21.log
