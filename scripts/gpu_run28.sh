cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
  for sk in 1 0; do for s in 128 64 256; do
  PARAKEET_B200_SPLITK=$sk timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --streams $s > gpurun_out/b28_$s.json 2> gpurun_out/b28_$s.err; echo "splitk=$sk streams=$s rc=$?"; tail -3 gpurun_out/b28_$s.err
  python -c "import json;d=json.load(open('gpurun_out/b28_$s.json'));print(d['ms_per_step'],d['value'])"
  done; done
) > gpurun_out/run28.log 2>&1
tail -20 gpurun_out/run28.log
