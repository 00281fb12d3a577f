cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency"
  timeout 600 python bench.py $BA > gpurun_out/b22s.json 2> gpurun_out/b22s.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 34400 -c 800 --csv --log-file gpurun_out/launches_r1d.csv python bench.py $BA > gpurun_out/b22_ncu1.log 2>&1
  echo "ncu1 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention -s 2210 -c 2 -o gpurun_out/prof_attn_r1d -f python bench.py $BA > gpurun_out/b22_ncu3.log 2>&1
  echo "ncu3 rc=$?"
) > gpurun_out/run22.log 2>&1
tail -10 gpurun_out/run22.log
