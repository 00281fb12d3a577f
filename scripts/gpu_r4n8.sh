# 2-GPU sanity of the sharded bench (torchrun, NCCL barrier + max-reduce only)
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r4n8_bench.json 2> gpurun_out/r4n8_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r4n8_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=8 ms/step', d['ms_per_step'], 'rtfx', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
