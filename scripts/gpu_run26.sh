cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
  for s in 1024 128; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency --streams $s > gpurun_out/b26_$s.json 2> gpurun_out/b26_$s.err; echo "streams=$s rc=$?"; tail -3 gpurun_out/b26_$s.err
  python -c "import json;d=json.load(open('gpurun_out/b26_$s.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'],d['roofline']['frac'])"
  done
) > gpurun_out/run26.log 2>&1
tail -14 gpurun_out/run26.log
