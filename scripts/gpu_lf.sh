# whole-utterance offline path: new tests first (fail fast), each test's outcome on its own line
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline_long.py tests/test_gpu_model.py::test_audio_mode_schedule_and_reset -m gpu -q -s -rA 2>&1 | tail -80 > gpurun_out/lf_tests.log
tail -60 gpurun_out/lf_tests.log
