cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( for d in 0 1; do
  if [ $d = 1 ]; then export PARAKEET_B200_DEFER=1; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b33_$d.json 2> gpurun_out/b33_$d.err; echo "defer=$d rc=$?"; tail -2 gpurun_out/b33_$d.err
  python -c "import json;d=json.load(open('gpurun_out/b33_$d.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'])"
  done
) > gpurun_out/run33.log 2>&1
cat gpurun_out/run33.log
