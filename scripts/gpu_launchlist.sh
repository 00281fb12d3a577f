cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${TAG:-tmp}
( BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency ${EXTRA:-}"
  timeout 600 python bench.py $BA > gpurun_out/ll_plain.json 2> gpurun_out/ll_plain.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s ${SKIP:-37000} -c 850 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $BA > gpurun_out/ll_ncu.log 2>&1
  echo "ncu rc=$?"
) > gpurun_out/ll.log 2>&1
tail -3 gpurun_out/ll.log
