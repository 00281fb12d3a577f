cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
  timeout 900 python bench.py > gpurun_out/b13.json 2> gpurun_out/b13.err; echo "bench rc=$?"; cat gpurun_out/b13.json; tail -3 gpurun_out/b13.err
  BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency"
  timeout 600 python bench.py $BA > gpurun_out/b13s.json 2> gpurun_out/b13s.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 34400 -c 800 --csv --log-file gpurun_out/launches_r1b.csv python bench.py $BA > gpurun_out/b13_ncu1.log 2>&1
  echo "ncu1 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 19200 -c 4 -o gpurun_out/prof_gemm_r1b -f python bench.py $BA > gpurun_out/b13_ncu2.log 2>&1
  echo "ncu2 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention -s 2210 -c 2 -o gpurun_out/prof_attn_r1b -f python bench.py $BA > gpurun_out/b13_ncu3.log 2>&1
  echo "ncu3 rc=$?"
) > gpurun_out/run13.log 2>&1
tail -30 gpurun_out/run13.log
