cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
( for shape in "768 1024 1024 f32" "768 1024 1024 resadd" "768 1024 4096 resadd" "768 4096 1024 silu" "768 3072 1024 f32" "768 2048 1024 glu" \
               "6144 1024 1024 resadd" "6144 1024 4096 resadd" "6144 4096 1024 silu" "6144 3072 1024 f32" "6144 2048 1024 glu" "128 1024 1024 f32" "384 1024 4096 resadd"; do
    set -- $shape
    timeout 60 $K gemm $1 $2 $3 100 $4 0 | tail -1
  done
  echo "--- with interleaved layernorm (smem carveout switches)"
  timeout 60 $K gemm 768 1024 1024 100 resadd 1 | tail -1
  timeout 60 $K gemm 6144 1024 1024 100 resadd 1 | tail -1
  echo "--- forced BN"
  for bn in 128 256; do for shape in "6144 4096 1024 silu" "6144 1024 4096 resadd" "768 4096 1024 silu"; do set -- $shape; PARAKEET_B200_GEMM_BN=$bn timeout 60 $K gemm $1 $2 $3 100 $4 0 | tail -1; done; done
) > gpurun_out/run19.log 2>&1
cat gpurun_out/run19.log
