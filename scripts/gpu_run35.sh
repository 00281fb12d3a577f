cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_frontend.py -m gpu -q -x 2>&1 | tail -2
  BA="--steps 2 --warmup 3 --no-cpu-baseline --no-latency --prefill-chunks 4"
  timeout 300 python bench.py $BA > gpurun_out/b35.json 2> gpurun_out/b35.err && \
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"logmel|subsample|layernorm|dwconv" -s 60 -c 30 --csv --log-file gpurun_out/misc35.csv python bench.py $BA > gpurun_out/b35_ncu.log 2>&1
  echo "ncu rc=$?"
) > gpurun_out/run35.log 2>&1
cat gpurun_out/run35.log
