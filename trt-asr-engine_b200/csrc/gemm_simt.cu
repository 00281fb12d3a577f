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
#include "enc_kernels.cuh"
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

// ------------------------------------------------------------------------------------------------ LayerNorm fused into the A operand
// One stream (M <= 16), bf16 mode: the three LayerNorms of a conformer layer that only feed ONE projection (norm_self_att -> q|k|v,
// norm_conv -> pointwise_conv1, norm_feed_forward2 -> FFN-up) were separate 6 us launches in a ~370-launch chain (29 % of the
// 1-stream chunk, profiles/r02_launch_summary_1stream.csv).  Here every CTA normalises its (<= 8) rows of the residual stream itself
// -- warp w owns row w, same lane <-> column mapping, summation order and bf16 rounding as layernorm_kernel, so the operand is
// bit-identical -- into a shared-memory copy of A; gamma / beta (constants) are staged before the dependency wait, like the weights.
// CTA 0 of a row group also writes the contract cache ring (norm_self_att) when that is enabled.  K == 1024.
template <int KS>
__global__ void __launch_bounds__(kWarps * 32, 2)
gemm_simt_ln_kernel(const GemmArgs g, const LnFuse f) {
  constexpr int kPairs = kWarps / KS;
  constexpr int kPitch = kDModel + 8;                       // bf16 elements per staged row (16-byte multiple, rows 16 B apart in banks)
  __shared__ float s_part[KS > 1 ? kWarps : 1][2 * kRows];
  __shared__ __align__(16) __nv_bfloat16 s_a[kRows][kPitch];
  __shared__ __align__(16) float s_gb[2][kDModel];
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
  {
    const int i = threadIdx.x;                              // 256 threads x float4 = 1024 floats
    reinterpret_cast<float4*>(s_gb[0])[i] = __ldg(reinterpret_cast<const float4*>(f.gamma) + i);
    reinterpret_cast<float4*>(s_gb[1])[i] = __ldg(reinterpret_cast<const float4*>(f.beta) + i);
  }
  __syncthreads();
  pdl_wait();
  const int M = g.M_dev ? min(*g.M_dev, g.M) : g.M;
  if (m0 + warp < M) {                                      // warp w: LayerNorm of row m0 + w -> bf16 operand row in shared memory
    const int row = m0 + warp;
    const float* xr = f.x + (size_t)row * kDModel;
    float v[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    ln_row(v, s_gb[0], s_gb[1], lane);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      store_act4(&s_a[warp][0], 0, 0, i * 128 + lane * 4, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), 0);
    if (f.has_ac && blockIdx.x == 0) {
      const AcacheOut& ac = f.ac;
      const int e = ac.row_entry[row];
      const int phys = (ac.entry_head[e] + kCacheS + ac.row_pos[row]) % kRingCap;
      const size_t base = ((size_t)ac.entry_slot[e] * kRingCap + phys) * kDModel;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = i * 128 + lane * 4;
        if (ac.is_f32) {
          *reinterpret_cast<float4*>((float*)ac.ring + base + col) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          store_act4((__nv_bfloat16*)ac.ring + base + col, 0, 0, 0, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), 0);
        }
      }
    }
  }
  __syncthreads();
  const bool live = n0 < g.N && m0 < M;
  float acc0[kRows], acc1[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r) acc0[r] = acc1[r] = 0.0f;
  if (live) {
    int row_off[kRows];                                     // element offsets of the (clamped) staged rows
#pragma unroll
    for (int r = 0; r < kRows; ++r) row_off[r] = min(r, M - 1 - m0) * kPitch;
    bool first = true;
    for (int k = kb + lane * 8; k < ke; k += 256 * kInflight) {
      if (!first) {
#pragma unroll
        for (int j = 0; j < kInflight; ++j) {
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
          unpack8(*reinterpret_cast<const uint4*>(&s_a[0][0] + row_off[r] + kk), af);
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
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    if (lane == r) { v0 = acc0[r]; v1 = acc1[r]; }
  if constexpr (KS > 1) {
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

bool gemm_simt_ln_supported(const GemmArgs& g) { return g.K == kDModel && g.a_lo_off == 0 && g.M > 0 && g.M <= 16 && g.N > 0; }

void gemm_simt_ln(const GemmArgs& g, const LnFuse& f, cudaStream_t st) {
  PKB_CHECK(gemm_simt_ln_supported(g), "gemm_simt_ln: K must be 1024, bf16 mode, M <= 16");
  int ks = g.N >= 2560 ? 1 : g.N >= 1536 ? 2 : 4;
  const int pairs = kWarps / ks;
  dim3 grid((g.N + 2 * pairs - 1) / (2 * pairs), (g.M + kRows - 1) / kRows);
  if (ks == 4) launch_k(gemm_simt_ln_kernel<4>, grid, dim3(kWarps * 32), 0, st, g, f);
  else if (ks == 2) launch_k(gemm_simt_ln_kernel<2>, grid, dim3(kWarps * 32), 0, st, g, f);
  else launch_k(gemm_simt_ln_kernel<1>, grid, dim3(kWarps * 32), 0, st, g, f);
  PKB_CUDA(cudaGetLastError());
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
