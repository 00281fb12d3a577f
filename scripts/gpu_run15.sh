cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -25
  echo "model rc=$?"
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
  BA="--steps 5 --warmup 3 --no-cpu-baseline --no-latency"
  timeout 600 python bench.py $BA > gpurun_out/b15.json 2> gpurun_out/b15.err; echo "bench rc=$?"; tail -3 gpurun_out/b15.err
  python -c "import json;d=json.load(open('gpurun_out/b15.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'],d['roofline']['frac'])"
) > gpurun_out/run15.log 2>&1
tail -50 gpurun_out/run15.log
