cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/b32_2gpu.json 2> gpurun_out/b32_2gpu.err; echo "rc=$?"; tail -5 gpurun_out/b32_2gpu.err; cat gpurun_out/b32_2gpu.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/b32_ref.json 2> gpurun_out/b32_ref.err; echo "ref rc=$?"; cat gpurun_out/b32_ref.json
) > gpurun_out/run32.log 2>&1
tail -12 gpurun_out/run32.log
