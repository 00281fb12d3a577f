cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
( nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 500 > gpurun_out/clk30.csv &
  SMI=$!
  for shape in "6144 4096 1024 silu" "6144 1024 4096 resadd" "6144 3072 1024 f32"; do set -- $shape; timeout 120 $K gemm $1 $2 $3 20000 $4 0 | tail -3; done
  kill $SMI
  sort gpurun_out/clk30.csv | uniq -c | sort -rn | head -8
) > gpurun_out/run30.log 2>&1
cat gpurun_out/run30.log
