cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
( for shape in "6144 4096 1024 silu" "6144 1024 4096 resadd" "6144 3072 1024 f32" "6144 1024 1024 resadd" "6144 2048 1024 glu"; do set -- $shape
    timeout 120 $K gemm $1 $2 $3 200 $4 0 1 | tail -1
    timeout 120 $K gemm $1 $2 $3 200 $4 0 40 | tail -1
  done
) > gpurun_out/run31.log 2>&1
cat gpurun_out/run31.log
