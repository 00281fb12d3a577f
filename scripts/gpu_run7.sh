cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -25
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b7.json 2> gpurun_out/b7.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 700 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b7_ncu.log 2>&1
  echo "ncu rc=$?"; cat gpurun_out/b7.json ) > gpurun_out/run7.log 2>&1
tail -40 gpurun_out/run7.log
