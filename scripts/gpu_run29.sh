cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency"
  timeout 600 python bench.py $BA > gpurun_out/b29s.json 2> gpurun_out/b29s.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 34400 -c 800 --csv --log-file gpurun_out/launches_r1f.csv python bench.py $BA > gpurun_out/b29_ncu1.log 2>&1
  echo "ncu1 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 19200 -c 10 -o gpurun_out/prof_gemm_r1f -f python bench.py $BA > gpurun_out/b29_ncu2.log 2>&1
  echo "ncu2 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention -s 2210 -c 2 -o gpurun_out/prof_attn_r1f -f python bench.py $BA > gpurun_out/b29_ncu3.log 2>&1
  echo "ncu3 rc=$?"
  timeout 900 ncu --set full --clock-control none -k regex:"logmel|tdt_select|lstm_cell|joint_hidden|layernorm|dwconv|subsample" -s 3000 -c 40 -o gpurun_out/prof_misc_r1f -f python bench.py $BA > gpurun_out/b29_ncu4.log 2>&1
  echo "ncu4 rc=$?"
) > gpurun_out/run29.log 2>&1
tail -10 gpurun_out/run29.log
