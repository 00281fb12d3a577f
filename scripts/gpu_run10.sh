cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|parity set|Error" | head -20
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b10.json 2> gpurun_out/b10.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 800 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b10_ncu.log 2>&1
  echo "ncu rc=$?"; cat gpurun_out/b10.json ) > gpurun_out/run10.log 2>&1
tail -40 gpurun_out/run10.log
