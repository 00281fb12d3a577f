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
          (2048, 4096, 1024), (1536, 3072, 1024), (700, 8198, 640), (3072, 1024, 4096)]


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
