cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
( echo "--- 2cta smoke"; PARAKEET_B200_GEMM_BN=512 timeout 30 $K gemm 512 512 256 5 f32 0 | tail -1; echo "rc=$?"
  if [ $? -eq 0 ]; then
  timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "2cta" 2>&1 | tail -6
  for shape in "6144 1024 1024 resadd" "6144 1024 4096 resadd" "6144 4096 1024 silu" "6144 3072 1024 f32" "6144 2048 1024 glu" "768 4096 1024 silu" "768 1024 4096 resadd" "1536 1024 4096 resadd"; do
    set -- $shape
    echo -n "1cta "; PARAKEET_B200_GEMM_2CTA=0 timeout 60 $K gemm $1 $2 $3 100 $4 0 | tail -1
    echo -n "2cta "; PARAKEET_B200_GEMM_BN=512 timeout 60 $K gemm $1 $2 $3 100 $4 0 | tail -1
  done
  fi
) > gpurun_out/run24.log 2>&1
cat gpurun_out/run24.log
