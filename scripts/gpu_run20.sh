cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
K=trt-asr-engine_b200/bin/kbench
( for shape in "768 1024 1024 resadd" "768 1024 4096 resadd" "768 4096 1024 silu" "768 3072 1024 f32" "768 2048 1024 glu" \
               "6144 1024 1024 resadd" "6144 1024 4096 resadd" "6144 4096 1024 silu" "6144 3072 1024 f32" "6144 2048 1024 glu"; do
    set -- $shape
    timeout 60 $K gemm $1 $2 $3 100 $4 0 | tail -1
  done
  for bn in 128 256; do for shape in "6144 4096 1024 silu" "6144 1024 4096 resadd" "6144 2048 1024 glu" "6144 1024 1024 resadd"; do set -- $shape; echo -n "bn=$bn "; PARAKEET_B200_GEMM_BN=$bn timeout 60 $K gemm $1 $2 $3 100 $4 0 | tail -1; done; done
  timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b20.json 2> gpurun_out/b20.err; echo "bench rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/b20.json'));print(d['ms_per_step'],d['value'],d['roofline']['achieved'],d['roofline']['frac'])"
) > gpurun_out/run20.log 2>&1
cat gpurun_out/run20.log
