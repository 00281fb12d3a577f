cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
  for s in ${STREAMS:-1024 128}; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --streams $s > gpurun_out/bq_$s.json 2> gpurun_out/bq_$s.err; echo "bench streams=$s rc=$?"; tail -3 gpurun_out/bq_$s.err
  python -c "import json;d=json.load(open('gpurun_out/bq_$s.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'],d['roofline']['frac'])"
  done
) > gpurun_out/quick.log 2>&1
tail -12 gpurun_out/quick.log
