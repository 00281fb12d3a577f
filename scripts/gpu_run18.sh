cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --streams 128"
  timeout 600 python bench.py $BA > gpurun_out/b18s.json 2> gpurun_out/b18s.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 34400 -c 800 --csv --log-file gpurun_out/launches_r1c_128.csv python bench.py $BA > gpurun_out/b18_ncu1.log 2>&1
  echo "ncu1 rc=$?"
) > gpurun_out/run18.log 2>&1
tail -3 gpurun_out/run18.log
