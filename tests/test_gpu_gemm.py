"""Both GEMM backends (weight-streaming CUDA-core kernel, tcgen05/TMA/TMEM kernel) against a float64 reference."""
import numpy as np
import pytest

import binding
from weights_io import bf16_bits_to_f32, f32_to_bf16_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[1, 0], ids=["precise", "bf16"])
def eng(request, model_small):
    e = binding.Engine(model_small, max_streams=1, precision=request.param)
    e.precision = request.param
    yield e
    e.close()


SHAPES = [(1, 256, 256), (6, 1024, 1024), (8, 4096, 1024), (17, 1024, 4096), (128, 256, 256), (130, 640, 640), (300, 1024, 1024),
          (257, 2560, 1280), (64, 8198, 640), (1000, 3072, 1024),
          # more tiles than SMs: the persistent loop takes >= 3 tiles per CTA, both TMEM accumulator buffers change phase
          (2048, 4096, 1024), (1536, 3072, 1024), (700, 8198, 640), (3072, 1024, 4096),
          # CTA-pair kernel with a partly filled last round: its tiles are cut into 2 (96 tiles on 74 pairs) or 4 column slices
          (6144, 1024, 1024), (4000, 2048, 1024)]


@pytest.mark.parametrize("backend", [0, 1, 2, 3, 4], ids=["simt", "tcgen05", "tcgen05-bn128", "tcgen05-bn256", "tcgen05-2cta"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm(eng, backend, M, N, K):
    rng = np.random.default_rng(M * 7 + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Wb = f32_to_bf16_bits(rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K))
    W = bf16_bits_to_f32(Wb).reshape(N, K)
    C = eng.gemm_test(backend, A, Wb.reshape(N, K))
    A_eff = A if eng.precision == 1 else bf16_bits_to_f32(f32_to_bf16_bits(A)).reshape(M, K)
    ref = A_eff.astype(np.float64) @ W.astype(np.float64).T
    err = np.max(np.abs(C - ref))
    # precise: hi+lo split carries 16 mantissa bits of A -> ~4e-6 relative per product; bf16: exact products, fp32 accumulation
    tol = 2e-4 if eng.precision == 1 else 1e-4
    assert err < tol * max(1.0, np.sqrt(K) / 16), (err, M, N, K)


def test_pair_kernel_tail_slices(model_small):
    """The optional column-slice tail of the CTA-pair kernel (PARAKEET_B200_GEMM_TAIL, off by default) in its own process -- the
    knob is read once per process: 96 tiles on 74 pairs cut in halves, 128 tiles cut in quarters, a ragged M."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import binding
from weights_io import bf16_bits_to_f32, f32_to_bf16_bits
for prec in (0, 1):
    e = binding.Engine(sys.argv[3], max_streams=1, precision=prec)
    for (M, N, K) in [(6144, 1024, 1024), (2048, 4096, 1024), (4000, 2048, 1024)]:
        rng = np.random.default_rng(M + N + K)
        A = rng.standard_normal((M, K)).astype(np.float32)
        Wb = f32_to_bf16_bits(rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K))
        W = bf16_bits_to_f32(Wb).reshape(N, K)
        C = e.gemm_test(4, A, Wb.reshape(N, K))
        A_eff = A if prec == 1 else bf16_bits_to_f32(f32_to_bf16_bits(A)).reshape(M, K)
        err = np.max(np.abs(C - A_eff.astype(np.float64) @ W.astype(np.float64).T))
        assert err < (2e-4 if prec == 1 else 1e-4) * max(1.0, np.sqrt(K) / 16), (err, M, N, K, prec)
    e.close()
print("TAIL-OK")
"""
    pkg = os.path.dirname(os.path.abspath(binding.__file__))
    env = dict(os.environ, PARAKEET_B200_GEMM_TAIL="4", PARAKEET_B200_GEMM_TAIL_T2="1", PARAKEET_B200_GEMM_TAIL_T4="1")
    out = subprocess.run([sys.executable, "-c", code, pkg, os.path.join(pkg, "tools"), model_small], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "TAIL-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
