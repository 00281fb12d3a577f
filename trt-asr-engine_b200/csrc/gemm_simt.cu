// gemm_simt.cu -- weight-streaming GEMM on CUDA cores (see gemm.h): the M <= 16 latency path (one or two streams).
//
// At one stream every projection is a matrix-vector-like product bound by the 1.2 GB of weights it streams from HBM, and every
// launch is small (2 - 17 MB): the whole weight matrix has to be IN FLIGHT at once for the launch to run at memory latency
// instead of a chain of dependent round trips.  Mapping: a warp owns two adjacent output columns (= two weight rows) over a
// k-slice; its lanes stride the slice with 16-byte loads and issue up to 8 of them (4 per column) before the first FMA; the
// KS warps that share a column pair (k-split inside the CTA, chosen so that ~2000 warps cover the matrix: KS = 4 for the
// N = 1024 projections, 2 for N = 2048, 1 above) add their partial sums through shared memory and the first of them runs the
// fused epilogue.  Activations (<= 8 rows per CTA row group, a few KB) come through L1.
// (r1's version -- one warp per column pair over the whole K, no k-split, one load per column in flight, 64 CTAs for an
// N = 1024 projection -- ran at 0.4 TB/s: 13.5 us per launch, 72 % of the 1-stream chunk latency; profiles/r02_launch_summary_1stream.csv.)
// It also serves as the reference implementation the tcgen05 path is checked against on the GPU (tests/test_gpu_gemm.py).
#include "gemm.h"

namespace pkb {

constexpr int kRows = 8;
constexpr int kWarps = 8;
constexpr int kInflight = 4;      // 16-byte weight loads per column a lane issues before it starts to multiply

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <int KS, bool SPLIT>
__global__ void __launch_bounds__(kWarps * 32, SPLIT ? 1 : 2)
gemm_simt_kernel(const GemmArgs g) {
  constexpr int kPairs = kWarps / KS;                       // column pairs per CTA
  __shared__ float s_part[KS > 1 ? kWarps : 1][2 * kRows];
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp / KS, ks = warp % KS;
  const int n0 = (blockIdx.x * kPairs + pair) * 2;
  const int m0 = blockIdx.y * kRows;
  const bool has1 = n0 + 1 < g.N;
  const int kslice = ((g.K + KS - 1) / KS + 7) & ~7;
  const int kb = ks * kslice, ke = min(g.K, kb + kslice);
  const __nv_bfloat16* w0 = g.W + (size_t)min(n0, g.N - 1) * g.K;
  const __nv_bfloat16* w1 = g.W + (size_t)(has1 ? n0 + 1 : min(n0, g.N - 1)) * g.K;
  // The weights are constants: the first batch of loads is requested BEFORE waiting for the predecessor kernel, so that in a chain
  // of dependent launches (one stream: ~370 per chunk) the HBM round trip of launch i+1 overlaps the execution of launch i.
  uint4 wa[kInflight], wb[kInflight];
  if (n0 < g.N) {
#pragma unroll
    for (int j = 0; j < kInflight; ++j) {
      const int kk = kb + lane * 8 + 256 * j;
      if (kk < ke) {
        wa[j] = __ldg(reinterpret_cast<const uint4*>(w0 + kk));
        wb[j] = __ldg(reinterpret_cast<const uint4*>(w1 + kk));
      }
    }
  }
  pdl_wait();
  const int M = g.M_dev ? min(*g.M_dev, g.M) : g.M;
  const bool live = n0 < g.N && m0 < M;                     // (dead warps still take part in the CTA barrier below)
  float acc0[kRows], acc1[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) acc0[r] = acc1[r] = 0.0f;
  if (live) {
    int row_off[kRows];                                     // element offsets of the (clamped) activation rows
#pragma unroll
    for (int r = 0; r < kRows; ++r) row_off[r] = min(m0 + r, M - 1) * g.lda;
    bool first = true;
    for (int k = kb + lane * 8; k < ke; k += 256 * kInflight) {
      if (!first) {
#pragma unroll
        for (int j = 0; j < kInflight; ++j) {               // every weight byte of this step is requested before any is consumed
          const int kk = k + 256 * j;
          if (kk < ke) {
            wa[j] = __ldg(reinterpret_cast<const uint4*>(w0 + kk));
            wb[j] = __ldg(reinterpret_cast<const uint4*>(w1 + kk));
          }
        }
      }
      first = false;
#pragma unroll
      for (int j = 0; j < kInflight; ++j) {
        const int kk = k + 256 * j;
        if (kk >= ke) break;
        float wf0[8], wf1[8];
        unpack8(wa[j], wf0);
        unpack8(wb[j], wf1);
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          float af[8];
          unpack8(*reinterpret_cast<const uint4*>(g.A + row_off[r] + kk), af);
          if constexpr (SPLIT) {
            float lf[8];
            unpack8(*reinterpret_cast<const uint4*>(g.A + g.a_lo_off + row_off[r] + kk), lf);
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
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      acc0[r] = warp_sum(acc0[r]);
      acc1[r] = warp_sum(acc1[r]);
    }
  }
  // lane r finishes row r
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    if (lane == r) { v0 = acc0[r]; v1 = acc1[r]; }
  if constexpr (KS > 1) {
    // the KS warps of a column pair add their k-slices in a fixed order (deterministic), the first one finishes
    if (lane < kRows) { s_part[warp][lane] = v0; s_part[warp][kRows + lane] = v1; }
    __syncthreads();
    if (ks != 0) return;
    v0 = 0.f; v1 = 0.f;
    if (lane < kRows) {
#pragma unroll
      for (int x = 0; x < KS; ++x) { v0 += s_part[warp + x][lane]; v1 += s_part[warp + x][kRows + lane]; }
    }
  }
  if (live && lane < kRows && m0 + lane < M) epilogue_pair(g.epi, m0 + lane, n0, g.N, v0, v1);
}

void gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  PKB_CHECK(g.K % 8 == 0 && g.lda % 8 == 0 && g.a_lo_off % 8 == 0, "gemm_simt: K, lda and a_lo_off must be multiples of 8");
  // k-split inside the CTA: enough warps in flight (~2000) to cover the matrix, at least 64 k-elements (8 lanes) per slice
  int ks = g.N >= 2560 ? 1 : g.N >= 1536 ? 2 : 4;
  while (ks > 1 && g.K / ks < 64) ks >>= 1;
  const int pairs = kWarps / ks;
  dim3 grid((g.N + 2 * pairs - 1) / (2 * pairs), (g.M + kRows - 1) / kRows);
  const bool split = g.a_lo_off != 0;
#define PKB_SIMT_CASE(KS_)                                                                      \
  if (split) launch_k(gemm_simt_kernel<KS_, true>, grid, dim3(kWarps * 32), 0, st, g);          \
  else launch_k(gemm_simt_kernel<KS_, false>, grid, dim3(kWarps * 32), 0, st, g);
  if (ks == 4) { PKB_SIMT_CASE(4) } else if (ks == 2) { PKB_SIMT_CASE(2) } else { PKB_SIMT_CASE(1) }
#undef PKB_SIMT_CASE
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
