// gemm_simt.cu -- weight-streaming GEMM on CUDA cores (see gemm.h).
//
// One warp owns two adjacent output columns (= two weight rows); its 32 lanes stride over K with 16-byte loads,
// so every weight byte is read exactly once per 8-row group, fully coalesced.  At M <= 16 (one or two streams)
// the projections are HBM-bound on the 1.2 GB of weights and this is the right shape; it also serves as the
// reference implementation the tcgen05 path is checked against on the GPU (tests/test_gemm_gpu.py).
#include "gemm.h"

namespace pkb {

constexpr int kRows = 8;
constexpr int kWarps = 8;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__global__ void __launch_bounds__(kWarps * 32)
gemm_simt_kernel(const GemmArgs g) {
  pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * kWarps + warp) * 2;
  const int m0 = blockIdx.y * kRows;
  const int M = g.M_dev ? min(*g.M_dev, g.M) : g.M;
  if (n0 >= g.N || m0 >= M) return;
  const bool split = g.a_lo_off != 0;
  const bool has1 = n0 + 1 < g.N;
  const __nv_bfloat16* w0 = g.W + (size_t)n0 * g.K;
  const __nv_bfloat16* w1 = g.W + (size_t)(has1 ? n0 + 1 : n0) * g.K;
  const __nv_bfloat16* arow[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) arow[r] = g.A + (size_t)min(m0 + r, M - 1) * g.lda;

  float acc0[kRows], acc1[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) acc0[r] = acc1[r] = 0.0f;

  for (int k = lane * 8; k < g.K; k += 256) {
    float wf0[8], wf1[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(w0 + k)), wf0);
    unpack8(__ldg(reinterpret_cast<const uint4*>(w1 + k)), wf1);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float af[8];
      unpack8(*reinterpret_cast<const uint4*>(arow[r] + k), af);
      if (split) {
        float lf[8];
        unpack8(*reinterpret_cast<const uint4*>(arow[r] + g.a_lo_off + k), lf);
#pragma unroll
        for (int i = 0; i < 8; ++i) af[i] += lf[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc0[r] = fmaf(af[i], wf0[i], acc0[r]);
        acc1[r] = fmaf(af[i], wf1[i], acc1[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    acc0[r] = warp_sum(acc0[r]);
    acc1[r] = warp_sum(acc1[r]);
  }
  // lane r finishes row r
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    if (lane == r) { v0 = acc0[r]; v1 = acc1[r]; }
  if (lane < kRows && m0 + lane < M) epilogue_pair(g.epi, m0 + lane, n0, g.N, v0, v1);
}

void gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  PKB_CHECK(g.K % 8 == 0 && g.lda % 8 == 0 && g.a_lo_off % 8 == 0, "gemm_simt: K, lda and a_lo_off must be multiples of 8");
  dim3 grid((g.N + 2 * kWarps - 1) / (2 * kWarps), (g.M + kRows - 1) / kRows);
  launch_k(gemm_simt_kernel, grid, dim3(kWarps * 32), 0, st, g);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
