# ncu captures of the default bench workload (1024 streams, 24 layers): launch list + full-metric captures of the hot kernels.
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
TAG=${TAG:-r1g}
( BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency"
  timeout 600 python bench.py $BA > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 34400 -c 800 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $BA > gpurun_out/prof_ncu1.log 2>&1
  echo "ncu1 rc=$?"
  timeout 900 ncu --set full --clock-control none -k regex:gemm_tc -s 19200 -c 9 -o gpurun_out/prof_gemm_$TAG -f python bench.py $BA > gpurun_out/prof_ncu2.log 2>&1
  echo "ncu2 rc=$?"
  timeout 900 ncu --set full --clock-control none -k regex:attention -s 2210 -c 1 -o gpurun_out/prof_attn_$TAG -f python bench.py $BA > gpurun_out/prof_ncu3.log 2>&1
  echo "ncu3 rc=$?"
  timeout 900 ncu --set full --clock-control none -k regex:"logmel|tdt_select|lstm_cell|joint_hidden|pred_input" -s 80 -c 12 -o gpurun_out/prof_misc_$TAG -f python bench.py $BA > gpurun_out/prof_ncu4.log 2>&1
  echo "ncu4 rc=$?"
  ls -la gpurun_out/*.ncu-rep
) > gpurun_out/profile.log 2>&1
tail -12 gpurun_out/profile.log
