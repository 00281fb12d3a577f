# round-2 GPU validation: new parity tests first (fail fast), then the whole GPU suite, then smoke
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q -s --durations=15 ${PYTEST_ARGS:-} > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -60 gpurun_out/r2_pytest.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
tail -5 gpurun_out/r2_smoke.log
