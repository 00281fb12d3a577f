// enc_kernels.cu -- see enc_kernels.cuh.
//
// Semantics restated from NeMo 2.6.0 (ConformerEncoder.forward_for_export, reached by the reference through
// /root/reference/tools/export_onnx/export.py:343-375; tensor contract contracts/parakeet-tdt-0.6b-v3.contract.json:97-159):
//   * ConvSubsampling dw_striding: conv0+ReLU, 2x(depthwise s2 + pointwise + ReLU), flatten (c*16+f), Linear
//   * RelPositionMultiHeadAttention over concat(cache, x); rel_shift == table row (Tk-Tq+i)-j; masked keys excluded
//   * CausalConv1D with a 4-column time cache, cache_drop_size 3
// B200-first differences from the exported graph: K/V are projected once and kept in per-stream ring buffers (the
// exported graph re-projects all 256 cached rows every chunk), linear_pos(pos_emb) is precomputed per layer,
// and the FIFO caches are rings addressed by a per-stream head, so carry-over costs O(new rows).
#include "enc_kernels.cuh"

namespace pkb {

__device__ __forceinline__ int find_entry(const int* __restrict__ prefix, int B, int idx) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (prefix[mid] <= idx) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------ rows
__global__ void build_rows_kernel(BatchDev b) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b.M) {
    const int e = find_entry(b.row_off, b.B, i);
    b.row_entry[i] = e;
    b.row_pos[i] = i - b.row_off[e];
  }
  if (i < b.sumT3) {
    const int e = find_entry(b.off3, b.B, i);
    const int t3 = i - b.off3[e];
    const int dp = b.drop[e];
    b.rowmap3[i] = (t3 >= dp && t3 - dp < b.Tq[e]) ? b.row_off[e] + t3 - dp : -1;
  }
}
void launch_build_rows(const BatchDev& b, cudaStream_t st) {
  const int n = b.M > b.sumT3 ? b.M : b.sumT3;
  if (n <= 0) return;
  launch_k(build_rows_kernel, dim3((n + 255) / 256), dim3(256), 0, st, b);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ subsampling stage 1
// One CTA per (entry, t2).  7 input rows -> conv0 (3 rows x 64 x 32 channels at a time, in smem) -> depthwise -> 32 x 256 outputs.
// Register tiling: a thread owns 2 adjacent channels and 4 adjacent conv0 columns, so the 9 inputs it loads per (row, tap row)
// feed 24 FMAs (the shared-memory loads, not the FMAs, bound the first version of this kernel).
__global__ void __launch_bounds__(256)
subsample_stage1_kernel(BatchDev b, const float* __restrict__ feat_ring, int ring_cap, SubsampleWeights w, ActOut a1) {
  pdl_enter();
  __shared__ float s_in[7][kNMels + 2];                 // [row][f+1], zero padded
  __shared__ __align__(8) float s_y0[3][66][34];        // [t1 row][f1+1][c]  (pitch 34: 8-byte aligned channel pairs)
  const int g = blockIdx.x;
  const int e = find_entry(b.off2, b.B, g);
  const int t2 = g - b.off2[e];
  const int T = b.T[e], T1 = b.T1[e];
  const float* ring = feat_ring + (size_t)b.slot[e] * ring_cap * kNMels;
  const int f0 = b.f0[e];
  const int tid = threadIdx.x;

  for (int i = tid; i < 7 * (kNMels + 2); i += 256) {
    const int r = i / (kNMels + 2), fp = i % (kNMels + 2);
    const int t = 4 * t2 - 3 + r, f = fp - 1;
    float v = 0.0f;
    if (t >= 0 && t < T && f >= 0 && f < kNMels) v = ring[(size_t)((f0 + t) % ring_cap) * kNMels + f];
    s_in[r][fp] = v;
  }
  const int cp = tid & 15, fgrp = tid >> 4;      // channel pair within the 32-channel group, 16 groups of 4 conv0 columns
  for (int cg = 0; cg < kSubCh / 32; ++cg) {
    const int c = cg * 32 + 2 * cp;
    float k0[2][9], k2[2][9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float2 ka = __ldg(reinterpret_cast<const float2*>(w.w0t + i * kSubCh + c));      // transposed [9][256] copies: channels c, c+1
      const float2 kb = __ldg(reinterpret_cast<const float2*>(w.w2t + i * kSubCh + c));
      k0[0][i] = ka.x; k0[1][i] = ka.y;
      k2[0][i] = kb.x; k2[1][i] = kb.y;
    }
    const float bias0[2] = {w.b0[c], w.b0[c + 1]}, bias2[2] = {w.b2[c], w.b2[c + 1]};
    __syncthreads();   // s_in ready (first iteration) / previous s_y0 consumers done
    // conv0 + ReLU for t1 = 2*t2-1 .. 2*t2+1, columns f1 = 4*fgrp .. 4*fgrp+3
#pragma unroll
    for (int r1 = 0; r1 < 3; ++r1) {
      const int t1 = 2 * t2 - 1 + r1;
      const bool row_ok = t1 >= 0 && t1 < T1;
      float acc[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[i][0] = bias0[0]; acc[i][1] = bias0[1]; }
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) {
        float in[9];                                   // padded inputs 8*fgrp .. 8*fgrp+8 of row 2*r1+dt
#pragma unroll
        for (int x = 0; x < 9; ++x) in[x] = s_in[2 * r1 + dt][8 * fgrp + x];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int df = 0; df < 3; ++df) {
            acc[i][0] = fmaf(k0[0][dt * 3 + df], in[2 * i + df], acc[i][0]);
            acc[i][1] = fmaf(k0[1][dt * 3 + df], in[2 * i + df], acc[i][1]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = row_ok ? make_float2(fmaxf(acc[i][0], 0.0f), fmaxf(acc[i][1], 0.0f)) : make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(&s_y0[r1][4 * fgrp + i + 1][2 * cp]) = v;
      }
      if (tid < 32) {                                   // f1 = -1 and f1 = 64 padding columns
        s_y0[r1][0][tid] = 0.0f;
        s_y0[r1][65][tid] = 0.0f;
      }
    }
    __syncthreads();
    // depthwise conv.2 (3x3, s2, p1): f2 = 2*fgrp, 2*fgrp+1
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f2 = 2 * fgrp + i;
      float a0 = bias2[0], a1v = bias2[1];
#pragma unroll
      for (int dt = 0; dt < 3; ++dt)
#pragma unroll
        for (int df = 0; df < 3; ++df) {
          const float2 y = *reinterpret_cast<const float2*>(&s_y0[dt][2 * f2 + df][2 * cp]);      // y0 col 2*f2-1+df == padded 2*f2+df
          a0 = fmaf(k2[0][dt * 3 + df], y.x, a0);
          a1v = fmaf(k2[1][dt * 3 + df], y.y, a1v);
        }
      store_act(a1.ptr, (size_t)g * 32 + f2, a1.lda, c, a0, a1.lo_off);
      store_act(a1.ptr, (size_t)g * 32 + f2, a1.lda, c + 1, a1v, a1.lo_off);
    }
  }
}
// ------------------------------------------------------------------------------------------------ stage 1 on tensor cores (bf16 mode)
// Same mapping (one CTA per (entry, t2)), but conv0 -- 192 positions x 256 channels x 9 taps -- runs as mma.sync m16n8k16
// (K = 9 taps padded to 16): A fragments are im2col patches built from the staged input rows, B fragments the bf16 filter
// taps, f32 accumulation, bias + ReLU on the accumulator fragments.  The CUDA-core version spends ~0.85 ms per 1024-stream
// step on this conv; here it is a few dozen MMAs per warp.  Channels are processed in two halves of 128 to keep the
// conv0 output tile (bf16 [3][66][128+8]) at 54 KB.
namespace {
constexpr int kS1Half = 128;                  // channels per pass
constexpr int kS1Pitch = kS1Half + 8;         // bf16 elements per (row, column) cell: 272 B -> conflict-free fragment stores
__device__ __forceinline__ uint32_t pack_bf16x2_f(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace

__global__ void __launch_bounds__(256, 3)
subsample_stage1_mma_kernel(BatchDev b, const float* __restrict__ feat_ring, int ring_cap, SubsampleWeights w, ActOut a1) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char s1_raw[];
  float (*s_in)[kNMels + 2] = reinterpret_cast<float (*)[kNMels + 2]>(s1_raw);                        // [7][130]
  __nv_bfloat16* s_y0 = reinterpret_cast<__nv_bfloat16*>(s1_raw + 7 * (kNMels + 2) * 4 + 8);          // [3][66][kS1Pitch]
  const int g = blockIdx.x;
  const int e = find_entry(b.off2, b.B, g);
  const int t2 = g - b.off2[e];
  const int T = b.T[e], T1 = b.T1[e];
  const float* ring = feat_ring + (size_t)b.slot[e] * ring_cap * kNMels;
  const int f0 = b.f0[e];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t = lane & 3;

  for (int i = tid; i < 7 * (kNMels + 2); i += 256) {
    const int r = i / (kNMels + 2), fp = i % (kNMels + 2);
    const int tt = 4 * t2 - 3 + r, f = fp - 1;
    float v = 0.0f;
    if (tt >= 0 && tt < T && f >= 0 && f < kNMels) v = ring[(size_t)((f0 + tt) % ring_cap) * kNMels + f];
    s_in[r][fp] = v;
  }
  // zero the f1 = -1 / 64 padding columns of the conv0 tile once (both passes reuse them)
  for (int i = tid; i < 3 * 2 * kS1Pitch; i += 256) {
    const int r1 = i / (2 * kS1Pitch), rest = i % (2 * kS1Pitch);
    s_y0[(r1 * 66 + (rest / kS1Pitch) * 65) * kS1Pitch + rest % kS1Pitch] = __float2bfloat16_rn(0.0f);
  }
  __syncthreads();

  // A fragments: m-tile mt covers row r1 = mt / 4, columns f1 = 16 (mt % 4) + {gq, gq + 8}; k index = tap 3 dt + df
  // lane t holds taps (2t, 2t+1) [a0: row gq, a1: row gq+8] and, for t == 0, tap 8 [a2, a3]
  uint32_t af[12][4];
  {
    const int k0 = 2 * t, k1 = 2 * t + 1;
    const int dt0 = k0 / 3, df0 = k0 % 3, dt1 = k1 / 3, df1 = k1 % 3;
#pragma unroll
    for (int mt = 0; mt < 12; ++mt) {
      const int r1 = mt >> 2, fb = 16 * (mt & 3);
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int f1 = fb + gq + 8 * hr;
        af[mt][hr] = pack_bf16x2_f(s_in[2 * r1 + dt0][2 * f1 + df0], s_in[2 * r1 + dt1][2 * f1 + df1]);
        af[mt][2 + hr] = t == 0 ? pack_bf16x2_f(s_in[2 * r1 + 2][2 * f1 + 2], 0.0f) : 0u;
      }
    }
  }

#pragma unroll 1
  for (int half = 0; half < kSubCh / kS1Half; ++half) {
    if (half) __syncthreads();      // the depthwise stage of the previous pass is done with s_y0
    // conv0 + bias + ReLU: warp w owns channels [half*128 + 16 w, +16) = 2 n-tiles
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int cl = 16 * warp + 8 * nt;                      // channel offset inside the half
      const int cB = half * kS1Half + cl + gq;                 // B fragment: channel = n index = gq
      const uint32_t b0 = pack_bf16x2_f(w.w0t[(2 * t) * kSubCh + cB], w.w0t[(2 * t + 1) * kSubCh + cB]);
      const uint32_t b1 = t == 0 ? pack_bf16x2_f(w.w0t[8 * kSubCh + cB], 0.0f) : 0u;
      const int cC = half * kS1Half + cl + 2 * t;              // C fragment: channels 2t, 2t+1
      const float bias0 = w.b0[cC], bias1 = w.b0[cC + 1];
#pragma unroll
      for (int mt = 0; mt < 12; ++mt) {
        float c[4] = {bias0, bias1, bias0, bias1};
        mma_bf16_16816(c, af[mt], b0, b1);
        const int r1 = mt >> 2, fb = 16 * (mt & 3);
        const bool row_ok = (2 * t2 - 1 + r1) >= 0 && (2 * t2 - 1 + r1) < T1;
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int f1 = fb + gq + 8 * hr;
          const uint32_t v = row_ok ? pack_bf16x2_f(fmaxf(c[2 * hr], 0.0f), fmaxf(c[2 * hr + 1], 0.0f)) : 0u;
          *reinterpret_cast<uint32_t*>(s_y0 + ((r1 * 66) + f1 + 1) * kS1Pitch + cl + 2 * t) = v;
        }
      }
    }
    __syncthreads();
    // depthwise conv.2 (3x3, s2, p1) on CUDA cores: thread = (channel pair cp of 64, quarter fq of the 32 output bins)
    {
      const int cp = tid & 63, fq = tid >> 6;
      const int c = half * kS1Half + 2 * cp;
      float k2[2][9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float2 kk = __ldg(reinterpret_cast<const float2*>(w.w2t + i * kSubCh + c));      // channels c, c+1 of tap i: 8 B per lane, coalesced
        k2[0][i] = kk.x; k2[1][i] = kk.y;
      }
      const float bias2[2] = {w.b2[c], w.b2[c + 1]};
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int f2 = 8 * fq + i;
        float a0 = bias2[0], a1v = bias2[1];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt)
#pragma unroll
          for (int df = 0; df < 3; ++df) {
            const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(s_y0 + ((dt * 66) + 2 * f2 + df) * kS1Pitch + 2 * cp);
            const float2 yf = __bfloat1622float2(y);
            a0 = fmaf(k2[0][dt * 3 + df], yf.x, a0);
            a1v = fmaf(k2[1][dt * 3 + df], yf.y, a1v);
          }
        const __nv_bfloat162 o = __floats2bfloat162_rn(a0, a1v);
        *reinterpret_cast<__nv_bfloat162*>(a1.ptr + ((size_t)g * 32 + f2) * a1.lda + c) = o;      // bf16 mode: hi plane only
      }
    }
  }
}

void launch_subsample_stage1(const BatchDev& b, const float* feat_ring, int ring_cap, const SubsampleWeights& w, ActOut a1,
                             cudaStream_t st) {
  if (b.sumT2 <= 0) return;
  static const bool use_mma = [] { const char* v = getenv("PARAKEET_B200_SUB_MMA"); return !(v && v[0] == '0'); }();
  if (a1.lo_off == 0 && use_mma) {      // bf16 mode: conv0 on tensor cores
    constexpr size_t smem = 7 * (kNMels + 2) * 4 + 8 + (size_t)3 * 66 * kS1Pitch * 2;
    static bool attr = false;
    if (!attr) {
      PKB_CUDA(cudaFuncSetAttribute(subsample_stage1_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    launch_k(subsample_stage1_mma_kernel, dim3(b.sumT2), dim3(256), smem, st, b, feat_ring, ring_cap, w, a1);
    return;
  }
  launch_k(subsample_stage1_kernel, dim3(b.sumT2), dim3(256), 0, st, b, feat_ring, ring_cap, w, a1);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ subsampling stage 2
// One CTA (256 threads) per (entry, t3): thread = (channel quad, output-bin quarter); 8-byte loads of 4 adjacent bf16 channels.
__global__ void __launch_bounds__(256)
subsample_stage2_kernel(BatchDev b, ActOut y1, SubsampleWeights w, ActOut a2) {
  pdl_enter();
  const int g = blockIdx.x;
  const int e = find_entry(b.off3, b.B, g);
  const int t3 = g - b.off3[e];
  const int T2 = b.T2[e];
  const int c = 4 * (threadIdx.x & 63), fq = threadIdx.x >> 6;      // channels c..c+3, output bins f3 = 4*fq .. 4*fq+3
  float k[4][9], bias[4];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float4 kk = __ldg(reinterpret_cast<const float4*>(w.w5t + i * kSubCh + c));      // channels c..c+3 of tap i: 16 B per lane, coalesced
    k[0][i] = kk.x; k[1][i] = kk.y; k[2][i] = kk.z; k[3][i] = kk.w;
  }
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(w.b5 + c));
    bias[0] = bb.x; bias[1] = bb.y; bias[2] = bb.z; bias[3] = bb.w;
  }
  const __nv_bfloat16* base = y1.ptr + (size_t)b.off2[e] * 32 * kSubCh;   // [T2][32][256]
#pragma unroll
  for (int i3 = 0; i3 < 4; ++i3) {
    const int f3 = 4 * fq + i3;
    float acc[4] = {bias[0], bias[1], bias[2], bias[3]};
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int t2 = 2 * t3 - 1 + dt;
      if (t2 < 0 || t2 >= T2) continue;
#pragma unroll
      for (int df = 0; df < 3; ++df) {
        const int f2 = 2 * f3 - 1 + df;
        if (f2 < 0 || f2 >= 32) continue;
        const size_t idx = ((size_t)t2 * 32 + f2) * kSubCh + c;
        const uint2 raw = *reinterpret_cast<const uint2*>(base + idx);
        float2 p0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        float2 p1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        if (y1.lo_off) {
          const uint2 lo = *reinterpret_cast<const uint2*>(base + idx + y1.lo_off);
          const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo.x));
          const float2 l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lo.y));
          p0.x += l0.x; p0.y += l0.y; p1.x += l1.x; p1.y += l1.y;
        }
        acc[0] = fmaf(k[0][dt * 3 + df], p0.x, acc[0]);
        acc[1] = fmaf(k[1][dt * 3 + df], p0.y, acc[1]);
        acc[2] = fmaf(k[2][dt * 3 + df], p1.x, acc[2]);
        acc[3] = fmaf(k[3][dt * 3 + df], p1.y, acc[3]);
      }
    }
    store_act4(a2.ptr, (size_t)g * 16 + f3, a2.lda, c, make_float4(acc[0], acc[1], acc[2], acc[3]), a2.lo_off);
  }
}
void launch_subsample_stage2(const BatchDev& b, ActOut y1, const SubsampleWeights& w, ActOut a2, cudaStream_t st) {
  if (b.sumT3 <= 0) return;
  launch_k(subsample_stage2_kernel, dim3(b.sumT3), dim3(256), 0, st, b, y1, w, a2);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// (ln_row: enc_kernels.cuh -- shared with the LayerNorm-fused CUDA-core GEMM)
constexpr int kLnRowsPerCta = 4;      // one warp per row; small CTAs balance 6144 rows over 148 SMs better than 8-row CTAs
// SPLITS = partial-sum planes of the preceding split-K GEMM, a compile-time count so that ALL their loads are in flight together with
// the x row (with a run-time loop over the planes each load waited for the previous one: 9.8 us per launch for 768 rows and three
// planes, 20 % of the 128-stream step, profiles/r02_launch_summary_128.csv)
template <int SPLITS>
__global__ void __launch_bounds__(kLnRowsPerCta * 32)
layernorm_kernel(float* __restrict__ x, int M, const float* __restrict__ g1, const float* __restrict__ b1,
                 const float* __restrict__ g2, const float* __restrict__ b2, int write_x, ActOut a, AcacheOut ac, int has_ac,
                 LnResidual res) {
  // gamma / beta are constants: staged in shared memory BEFORE the dependency wait (in a chain of small dependent launches their
  // round trip used to sit between the row statistics and the stores)
  __shared__ __align__(16) float s_gb[4][kDModel];
  pdl_trigger();
  for (int i = threadIdx.x; i < kDModel / 4; i += kLnRowsPerCta * 32) {
    reinterpret_cast<float4*>(s_gb[0])[i] = __ldg(reinterpret_cast<const float4*>(g1) + i);
    reinterpret_cast<float4*>(s_gb[1])[i] = __ldg(reinterpret_cast<const float4*>(b1) + i);
    if (g2 != nullptr) {
      reinterpret_cast<float4*>(s_gb[2])[i] = __ldg(reinterpret_cast<const float4*>(g2) + i);
      reinterpret_cast<float4*>(s_gb[3])[i] = __ldg(reinterpret_cast<const float4*>(b2) + i);
    }
  }
  __syncthreads();
  pdl_wait();
  const int row = blockIdx.x * kLnRowsPerCta + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  g1 = s_gb[0]; b1 = s_gb[1];
  if (g2 != nullptr) { g2 = s_gb[2]; b2 = s_gb[3]; }
  float v[32];
  float* xr = x + (size_t)row * kDModel;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(xr + i * 128 + lane * 4);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  if constexpr (SPLITS > 0) {
    // deferred residual of the preceding split-K GEMM: x += scale * (p_0 + p_1 + ...), splits added in index order
    if (res.bf16) {
      uint2 raw[SPLITS][8];
#pragma unroll
      for (int sp = 0; sp < SPLITS; ++sp)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          raw[sp][i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(res.part) + (size_t)sp * res.split_stride +
                                                       (size_t)row * kDModel + i * 128 + lane * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int sp = 0; sp < SPLITS; ++sp) {
          const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[sp][i].x));
          const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[sp][i].y));
          acc.x += lo.x; acc.y += lo.y; acc.z += hi.x; acc.w += hi.y;
        }
        v[4 * i] += res.scale * acc.x; v[4 * i + 1] += res.scale * acc.y; v[4 * i + 2] += res.scale * acc.z; v[4 * i + 3] += res.scale * acc.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 t[SPLITS];
#pragma unroll
        for (int sp = 0; sp < SPLITS; ++sp)
          t[sp] = *reinterpret_cast<const float4*>(res.part + (size_t)sp * res.split_stride + (size_t)row * kDModel + i * 128 + lane * 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int sp = 0; sp < SPLITS; ++sp) { acc.x += t[sp].x; acc.y += t[sp].y; acc.z += t[sp].z; acc.w += t[sp].w; }
        v[4 * i] += res.scale * acc.x; v[4 * i + 1] += res.scale * acc.y; v[4 * i + 2] += res.scale * acc.z; v[4 * i + 3] += res.scale * acc.w;
      }
    }
    if (!write_x) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(xr + i * 128 + lane * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }
  ln_row(v, g1, b1, lane);
  if (has_ac) {
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
  if (write_x) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(xr + i * 128 + lane * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    if (g2 != nullptr) ln_row(v, g2, b2, lane);
  }
  if (a.ptr == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    store_act4(a.ptr, row, a.lda, i * 128 + lane * 4, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), a.lo_off);
}
// ---- streaming LayerNorm for large batches (same arithmetic, lane <-> column mapping and summation order as layernorm_kernel: the
// results are bit-identical; only the data movement differs).  layernorm_kernel keeps 24 warps per SM resident, each with ONE row
// (4 KB of x + the partial sums) in flight between dependent phases: 1.6 TB/s of DRAM reads at 1024 streams (profiles/r02_misc...:
// 38 - 50 MB in 23 - 29 us, 12 % of the step).  Here one persistent CTA per SM owns <= 16 warps; a warp walks rows
// gw, gw + stride, ... and keeps `depth` rows in flight through cp.async.bulk copies (x row + partial-sum rows -> its private
// shared-memory ring, one mbarrier per slot), so the loads of row k+depth overlap the arithmetic and stores of row k; gamma / beta
// are staged in shared memory once per CTA, before the dependency wait.
namespace {
__device__ __forceinline__ uint32_t lns_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lns_bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lns_u32(dst)), "l"(src),
               "r"(bytes), "r"(lns_u32(bar))
               : "memory");
}
__device__ __forceinline__ void lns_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(lns_u32(bar)), "r"(parity)
        : "memory");
  }
}
constexpr int kLnsParamBytes = 4 * kDModel * 4;      // gamma1 | beta1 | gamma2 | beta2
constexpr int kLnsBarBytes = 1024;
}  // namespace

template <int SPLITS, bool PBF16>      // partial-sum planes of the preceding split-K GEMM (0, 1 or 2), their element type
__global__ void __launch_bounds__(512, 1)
layernorm_stream_kernel(float* __restrict__ x, int M, const float* __restrict__ g1, const float* __restrict__ b1,
                        const float* __restrict__ g2, const float* __restrict__ b2, int write_x, ActOut a, AcacheOut ac, int has_ac,
                        LnResidual res, int depth, int row_bytes) {
  extern __shared__ __align__(128) uint8_t lns_raw[];
  float* s_g1 = reinterpret_cast<float*>(lns_raw);
  float* s_b1 = s_g1 + kDModel;
  float* s_g2 = s_b1 + kDModel;
  float* s_b2 = s_g2 + kDModel;
  uint64_t* bars = reinterpret_cast<uint64_t*>(lns_raw + kLnsParamBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* mybuf = lns_raw + kLnsParamBytes + kLnsBarBytes + (size_t)warp * depth * row_bytes;
  uint64_t* mybar = bars + warp * depth;
  pdl_trigger();
  for (int i = threadIdx.x; i < kDModel / 4; i += blockDim.x) {      // constants: staged before the dependency wait
    reinterpret_cast<float4*>(s_g1)[i] = reinterpret_cast<const float4*>(g1)[i];
    reinterpret_cast<float4*>(s_b1)[i] = reinterpret_cast<const float4*>(b1)[i];
    if (g2 != nullptr) {
      reinterpret_cast<float4*>(s_g2)[i] = reinterpret_cast<const float4*>(g2)[i];
      reinterpret_cast<float4*>(s_b2)[i] = reinterpret_cast<const float4*>(b2)[i];
    }
  }
  if (lane == 0) {
    for (int d = 0; d < depth; ++d)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lns_u32(&mybar[d])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  const int gw = warp * gridDim.x + blockIdx.x, stride = nw * gridDim.x;
  constexpr uint32_t part_bytes = SPLITS > 0 ? (PBF16 ? kDModel * 2u : kDModel * 4u) : 0u;
  constexpr int splits = SPLITS;
  auto issue = [&](int row, int slot) {      // one lane: x row + the row of every partial-sum plane into ring slot `slot`
    uint8_t* dst = mybuf + (size_t)slot * row_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lns_u32(&mybar[slot])), "r"(kDModel * 4u + splits * part_bytes)
                 : "memory");
    lns_bulk(dst, x + (size_t)row * kDModel, kDModel * 4u, &mybar[slot]);
#pragma unroll
    for (int sp = 0; sp < splits; ++sp) {
      const size_t off = (size_t)sp * res.split_stride + (size_t)row * kDModel;
      const void* src = PBF16 ? static_cast<const void*>(reinterpret_cast<const __nv_bfloat16*>(res.part) + off) : static_cast<const void*>(res.part + off);
      lns_bulk(dst + kDModel * 4 + (size_t)sp * part_bytes, src, part_bytes, &mybar[slot]);
    }
  };
  if (lane == 0) {
    for (int d = 0; d < depth; ++d) {
      const int row = gw + d * stride;
      if (row < M) issue(row, d);
    }
  }
  int k = 0;
  for (int row = gw; row < M; row += stride, ++k) {
    const int slot = k % depth;
    lns_wait(&mybar[slot], (uint32_t)((k / depth) & 1));
    const uint8_t* buf = mybuf + (size_t)slot * row_bytes;
    float v[32];
    float* xr = x + (size_t)row * kDModel;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = reinterpret_cast<const float4*>(buf)[i * 32 + lane];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    if constexpr (SPLITS > 0) {
      // x += scale * (p_0 + p_1): the same sums, in the same order, as layernorm_kernel (0 + p_0 is exact), on the packed f32x2 pipe
      const float2 sc2 = make_float2(res.scale, res.scale);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int sp = 0; sp < SPLITS; ++sp) {
          const uint8_t* pb = buf + kDModel * 4 + (size_t)sp * part_bytes;
          float2 t0, t1;
          if constexpr (PBF16) {
            const uint2 raw = reinterpret_cast<const uint2*>(pb)[i * 32 + lane];
            t0 = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
            t1 = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
          } else {
            const float4 t = reinterpret_cast<const float4*>(pb)[i * 32 + lane];
            t0 = make_float2(t.x, t.y); t1 = make_float2(t.z, t.w);
          }
          if (sp == 0) { a0 = t0; a1 = t1; }
          else { a0 = __fadd2_rn(a0, t0); a1 = __fadd2_rn(a1, t1); }
        }
        const float2 r0 = __ffma2_rn(sc2, a0, make_float2(v[4 * i], v[4 * i + 1]));
        const float2 r1 = __ffma2_rn(sc2, a1, make_float2(v[4 * i + 2], v[4 * i + 3]));
        v[4 * i] = r0.x; v[4 * i + 1] = r0.y; v[4 * i + 2] = r1.x; v[4 * i + 3] = r1.y;
      }
    }
    // every lane holds its share of the row in registers: the slot can take the row `depth` iterations ahead
    __syncwarp();
    if (lane == 0) {
      const int next = row + depth * stride;
      if (next < M) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(next, slot);
      }
    }
    if (SPLITS > 0 && !write_x) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(xr + i * 128 + lane * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
    ln_row(v, s_g1, s_b1, lane);
    if (has_ac) {
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
    if (write_x) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(xr + i * 128 + lane * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (g2 != nullptr) ln_row(v, s_g2, s_b2, lane);
    }
    if (a.ptr != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        store_act4(a.ptr, row, a.lda, i * 128 + lane * 4, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]), a.lo_off);
    }
  }
}

void launch_layernorm(float* x, int M, const float* g1, const float* b1, const float* g2, const float* b2, int write_x, ActOut a,
                      const AcacheOut* ac, cudaStream_t st, const LnResidual* res) {
  if (M <= 0) return;
  AcacheOut z{};
  LnResidual r0{};
  // large batches: the streaming kernel (PARAKEET_B200_LN_STREAM=0 keeps layernorm_kernel everywhere; the two are bit-identical)
  static const int stream_min_rows = [] { const char* v = getenv("PARAKEET_B200_LN_STREAM"); return v ? (atoi(v) > 0 ? atoi(v) : 1 << 30) : 2048; }();
  if (M >= stream_min_rows) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      PKB_CUDA(cudaGetDevice(&dev));
      PKB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int splits = (res && res->part) ? res->splits : 0;
    const bool pbf16 = res && res->bf16;
    const int row_bytes = kDModel * 4 + splits * (pbf16 ? kDModel * 2 : kDModel * 4);
    // 16 warps x 2 rows in flight where the rows are small enough (bf16 mode, one partial-sum plane: 6 KB per row), else 8 warps x 3:
    // with 8 warps the kernel is bound by the latency of its own instruction chain (2 warps per scheduler: 23 us at 1024 streams,
    // profiles/r02_ln_stream...), not by memory
    static const int budget_kb = [] { const char* v = getenv("PARAKEET_B200_LN_SMEM_KB"); return v ? atoi(v) : 210; }();
    const int kBudget = budget_kb * 1024;
    int nw = 16, depth = kBudget / (nw * row_bytes);
    if (depth < 2) { nw = 12; depth = kBudget / (nw * row_bytes); }      // two bf16 partial-sum planes: 8 KB per row
    if (depth < 2) { nw = 8; depth = kBudget / (nw * row_bytes); }
    if (depth >= 2 && splits <= 2) {
      depth = depth > 3 ? 3 : depth;
      const size_t smem = kLnsParamBytes + kLnsBarBytes + (size_t)nw * depth * row_bytes;
      const int grid = (M + nw - 1) / nw < sms ? (M + nw - 1) / nw : sms;
      auto go = [&](auto kernel) {
        PKB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        launch_k(kernel, dim3(grid), dim3(nw * 32), smem, st, x, M, g1, b1, g2, b2, write_x, a, ac ? *ac : z, (int)(ac != nullptr), res ? *res : r0, depth,
                 row_bytes);
      };
      if (splits == 0) go(layernorm_stream_kernel<0, false>);
      else if (splits == 1 && pbf16) go(layernorm_stream_kernel<1, true>);
      else if (splits == 2 && pbf16) go(layernorm_stream_kernel<2, true>);
      else if (splits == 1) go(layernorm_stream_kernel<1, false>);
      else go(layernorm_stream_kernel<2, false>);
      PKB_CUDA(cudaGetLastError());
      return;
    }
  }
  const int nsp = (res && res->part) ? res->splits : 0;
  PKB_CHECK(nsp >= 0 && nsp <= 4, "layernorm: at most 4 partial-sum planes");
  auto go = [&](auto kernel) {
    launch_k(kernel, dim3((M + kLnRowsPerCta - 1) / kLnRowsPerCta), dim3(kLnRowsPerCta * 32), 0, st, x, M, g1, b1, g2, b2, write_x, a, ac ? *ac : z,
             (int)(ac != nullptr), res ? *res : r0);
  };
  switch (nsp) {
    case 0: go(layernorm_kernel<0>); break;
    case 1: go(layernorm_kernel<1>); break;
    case 2: go(layernorm_kernel<2>); break;
    case 3: go(layernorm_kernel<3>); break;
    default: go(layernorm_kernel<4>); break;
  }
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ attention
// One CTA (128 threads) per (entry, head).
//   phase A1: thread == key column j (coalesced over the K^T ring):  ac[i][j] = (q_i + u) . k_j
//   phase A2: thread == relative-position column:                      G[i][r]  = (q_i + v) . p_r
//   phase B : scores[i][j] = (ac[i][j] + G[i][(256+i-j)+kPosNeg]) / sqrt(128), masked softmax (one warp per query row)
//   phase C : thread == head dim d (coalesced over the V ring rows):   ctx[i][d] = sum_j p[i][j] v_j[d]
constexpr int kQG = 6;   // queries processed together in registers (steady-state Tq)

template <typename KV> __device__ __forceinline__ float ld_kv(const KV* p);
template <> __device__ __forceinline__ float ld_kv<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_kv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename KV>
__global__ void __launch_bounds__(128)
attention_kernel(BatchDev b, AttnArgs a) {
  pdl_enter();
  extern __shared__ float smem[];
  const int e = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
  const int Tq = b.Tq[e], qlen = b.qlen[e], len = b.len[e], head = b.head[e], slot = b.slot[e];
  const int row0 = b.row_off[e];
  const int Tk = kCacheS + Tq;
  const int npos = Tk + Tq - 1;                 // relative positions -(Tq-1) .. Tk-1
  float* s_qu = smem;                           // [Tq][128]
  float* s_qv = s_qu + Tq * kDHead;             // [Tq][128]
  float* s_sc = s_qv + Tq * kDHead;             // [Tq][kRingCap]   scores / probabilities
  float* s_g = s_sc + Tq * kRingCap;            // [Tq][kPosRows]   position scores

  for (int i = tid; i < Tq * kDHead; i += 128) {
    const int qi = i >> 7, d = i & 127;
    const float qv = a.q[(size_t)(row0 + qi) * kDModel + h * kDHead + d];
    s_qu[i] = qv + a.bias_u[h * kDHead + d];
    s_qv[i] = qv + a.bias_v[h * kDHead + d];
  }
  __syncthreads();

  const KV* kt = reinterpret_cast<const KV*>(a.kring) + ((size_t)slot * kHeads + h) * kDHead * kRingCap;
  const KV* pt = reinterpret_cast<const KV*>(a.ppos_t) + (size_t)h * kDHead * kPosRows;
  const int r_lo = kPosNeg - (Tq - 1);          // first table row used
  for (int q0 = 0; q0 < Tq; q0 += kQG) {
    const int nq = min(kQG, Tq - q0);
    for (int j = kCacheS - len + tid; j < kCacheS + qlen; j += 128) {       // A1 (valid keys only)
      const int phys = (head + j) % kRingCap;
      float acc[kQG];
#pragma unroll
      for (int i = 0; i < kQG; ++i) acc[i] = 0.0f;
      for (int d = 0; d < kDHead; ++d) {
        const float kv = ld_kv<KV>(kt + (size_t)d * kRingCap + phys);
#pragma unroll
        for (int i = 0; i < kQG; ++i)
          if (i < nq) acc[i] = fmaf(s_qu[(q0 + i) * kDHead + d], kv, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < kQG; ++i)
        if (i < nq) s_sc[(q0 + i) * kRingCap + j] = acc[i];
    }
    for (int c = tid; c < npos; c += 128) {     // A2
      const int r = r_lo + c;
      float acc[kQG];
#pragma unroll
      for (int i = 0; i < kQG; ++i) acc[i] = 0.0f;
      for (int d = 0; d < kDHead; ++d) {
        const float pv = ld_kv<KV>(pt + (size_t)d * kPosRows + r);
#pragma unroll
        for (int i = 0; i < kQG; ++i)
          if (i < nq) acc[i] = fmaf(s_qv[(q0 + i) * kDHead + d], pv, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < kQG; ++i)
        if (i < nq) s_g[(q0 + i) * kPosRows + r] = acc[i];
    }
  }
  __syncthreads();

  // phase B: masked softmax; key j valid iff (j >= 256-len for cached rows) or (j-256 < qlen for new rows)
  const float scale = 0.08838834764831845f;     // 1/sqrt(128)
  const int warp = tid >> 5, lane = tid & 31;
  const int j_lo = kCacheS - len, j_hi = kCacheS + qlen;
  for (int i = warp; i < Tq; i += 4) {
    float mx = -INFINITY;
    for (int j = j_lo + lane; j < j_hi; j += 32) {
      const float s = (s_sc[i * kRingCap + j] + s_g[i * kPosRows + (kCacheS + i - j) + kPosNeg]) * scale;
      s_sc[i * kRingCap + j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = j_lo + lane; j < j_hi; j += 32) {
      const float p = __expf(s_sc[i * kRingCap + j] - mx);
      s_sc[i * kRingCap + j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    const float inv = (i < qlen && sum > 0.0f) ? 1.0f / sum : 0.0f;   // padded query rows are fully masked -> zeros
    for (int j = j_lo + lane; j < j_hi; j += 32) s_sc[i * kRingCap + j] *= inv;
  }
  __syncthreads();

  // phase C
  const KV* vr = reinterpret_cast<const KV*>(a.vring) + (size_t)slot * kRingCap * kDModel + h * kDHead + tid;
  for (int q0 = 0; q0 < Tq; q0 += kQG) {
    const int nq = min(kQG, Tq - q0);
    float acc[kQG];
#pragma unroll
    for (int i = 0; i < kQG; ++i) acc[i] = 0.0f;
    for (int j = j_lo; j < j_hi; ++j) {
      const int phys = (head + j) % kRingCap;
      const float vv = ld_kv<KV>(vr + (size_t)phys * kDModel);
#pragma unroll
      for (int i = 0; i < kQG; ++i)
        if (i < nq) acc[i] = fmaf(s_sc[(q0 + i) * kRingCap + j], vv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < kQG; ++i)
      if (i < nq) store_act(a.ctx.ptr, row0 + q0 + i, a.ctx.lda, h * kDHead + tid, acc[i], a.ctx.lo_off);
  }
}
void launch_attention(const BatchDev& b, const AttnArgs& a, cudaStream_t st) {
  if (b.B <= 0) return;
  const size_t smem = (size_t)b.max_Tq * (2 * kDHead + kRingCap + kPosRows) * sizeof(float);
  static size_t attr_set[2] = {0, 0};
  if (smem > 48 * 1024 && attr_set[a.kv_f32] < smem) {
    if (a.kv_f32) PKB_CUDA(cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else PKB_CUDA(cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[a.kv_f32] = smem;
  }
  if (a.kv_f32) launch_k(attention_kernel<float>, dim3(b.B, kHeads), dim3(128), smem, st, b, a);
  else launch_k(attention_kernel<__nv_bfloat16>, dim3(b.B, kHeads), dim3(128), smem, st, b, a);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ conv module middle
// One CTA per (entry, 256-channel group); thread == channel.
__global__ void __launch_bounds__(256)
dwconv_kernel(BatchDev b, DwConvArgs a) {
  pdl_enter();
  const int e = blockIdx.x, ch = blockIdx.y * 256 + threadIdx.x;
  const int Tq = b.Tq[e], qlen = b.qlen[e], row0 = b.row_off[e];
  float* cache = a.cache_tm + (size_t)b.slot[e] * a.slot_stride + (size_t)ch * kTimeCtx;
  float w[kConvK];
#pragma unroll
  for (int i = 0; i < kConvK; ++i) w[i] = a.w[ch * kConvK + i];
  const float bias = a.bias[ch];
  auto ld_c = [&](int t) -> float {
    const size_t i = (size_t)(row0 + t) * kDModel + ch;
    return a.c_bf16 ? __bfloat162float(a.c_bf16[i]) : a.c[i];
  };
  // ext = [cache(4) | c(Tq, zero past qlen: pad_mask) | 0 0 0 0]; sliding window of 9
  const bool offline = b.offline[e] != 0;      // offline: zero left context (symmetric 4/4 padding), cache untouched
  const float4 cv = offline ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(cache);
  float win[kConvK];
  win[0] = cv.x; win[1] = cv.y; win[2] = cv.z; win[3] = cv.w;
#pragma unroll
  for (int i = 4; i < kConvK; ++i) {
    const int t = i - 4;
    win[i] = (t < Tq && t < qlen) ? ld_c(t) : 0.0f;
  }
  // new cache = ext[Tq+1 .. Tq+5)   (new_x[:-3][-4:], cache_drop_size 3)
  float nc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = Tq + 1 + i;                 // index into ext
    float v = 0.0f;
    if (idx < 4) v = idx == 0 ? cv.x : idx == 1 ? cv.y : idx == 2 ? cv.z : cv.w;
    else if (idx - 4 < Tq) v = (idx - 4 < qlen) ? ld_c(idx - 4) : 0.0f;
    nc[i] = v;
  }
  for (int t = 0; t < Tq; ++t) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < kConvK; ++i) acc = fmaf(w[i], win[i], acc);
    acc += bias;
    store_act(a.out.ptr, row0 + t, a.out.lda, ch, silu(acc), a.out.lo_off);
#pragma unroll
    for (int i = 0; i < kConvK - 1; ++i) win[i] = win[i + 1];
    const int tn = t + 5;                       // next ext index t+1+8 -> c index t+5
    win[kConvK - 1] = (tn < Tq && tn < qlen) ? ld_c(tn) : 0.0f;
  }
  if (!offline) *reinterpret_cast<float4*>(cache) = make_float4(nc[0], nc[1], nc[2], nc[3]);
}
// bf16 mode: two adjacent channels per thread (4-byte loads / stores of bf16 pairs: 128 B per warp and row instead of 64 B, half
// the threads); per channel the arithmetic and its order are exactly those of dwconv_kernel.
__global__ void __launch_bounds__(256)
dwconv2_kernel(BatchDev b, DwConvArgs a) {
  pdl_enter();
  const int e = blockIdx.x, ch = 2 * (blockIdx.y * 256 + threadIdx.x);
  const int Tq = b.Tq[e], qlen = b.qlen[e], row0 = b.row_off[e];
  float* cache = a.cache_tm + (size_t)b.slot[e] * a.slot_stride + (size_t)ch * kTimeCtx;
  float w[2][kConvK];
#pragma unroll
  for (int i = 0; i < kConvK; ++i) { w[0][i] = a.w[ch * kConvK + i]; w[1][i] = a.w[(ch + 1) * kConvK + i]; }
  const float2 bias = *reinterpret_cast<const float2*>(a.bias + ch);
  auto ld_c = [&](int t) -> float2 {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a.c_bf16 + (size_t)(row0 + t) * kDModel + ch));
  };
  const bool offline = b.offline[e] != 0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 c0 = offline ? z4 : *reinterpret_cast<const float4*>(cache);
  const float4 c1 = offline ? z4 : *reinterpret_cast<const float4*>(cache + kTimeCtx);
  float win[2][kConvK];
  win[0][0] = c0.x; win[0][1] = c0.y; win[0][2] = c0.z; win[0][3] = c0.w;
  win[1][0] = c1.x; win[1][1] = c1.y; win[1][2] = c1.z; win[1][3] = c1.w;
#pragma unroll
  for (int i = 4; i < kConvK; ++i) {
    const int t = i - 4;
    const float2 v = (t < Tq && t < qlen) ? ld_c(t) : make_float2(0.f, 0.f);
    win[0][i] = v.x; win[1][i] = v.y;
  }
  float nc[2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = Tq + 1 + i;
    float2 v = make_float2(0.f, 0.f);
    if (idx < 4) v = idx == 0 ? make_float2(c0.x, c1.x) : idx == 1 ? make_float2(c0.y, c1.y) : idx == 2 ? make_float2(c0.z, c1.z) : make_float2(c0.w, c1.w);
    else if (idx - 4 < Tq) v = (idx - 4 < qlen) ? ld_c(idx - 4) : make_float2(0.f, 0.f);
    nc[0][i] = v.x; nc[1][i] = v.y;
  }
  for (int t = 0; t < Tq; ++t) {
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
    for (int i = 0; i < kConvK; ++i) { acc0 = fmaf(w[0][i], win[0][i], acc0); acc1 = fmaf(w[1][i], win[1][i], acc1); }
    acc0 += bias.x; acc1 += bias.y;
    *reinterpret_cast<__nv_bfloat162*>(a.out.ptr + (size_t)(row0 + t) * a.out.lda + ch) = __floats2bfloat162_rn(silu(acc0), silu(acc1));
#pragma unroll
    for (int i = 0; i < kConvK - 1; ++i) { win[0][i] = win[0][i + 1]; win[1][i] = win[1][i + 1]; }
    const int tn = t + 5;
    const float2 v = (tn < Tq && tn < qlen) ? ld_c(tn) : make_float2(0.f, 0.f);
    win[0][kConvK - 1] = v.x; win[1][kConvK - 1] = v.y;
  }
  if (!offline) {
    *reinterpret_cast<float4*>(cache) = make_float4(nc[0][0], nc[0][1], nc[0][2], nc[0][3]);
    *reinterpret_cast<float4*>(cache + kTimeCtx) = make_float4(nc[1][0], nc[1][1], nc[1][2], nc[1][3]);
  }
}
// bf16 mode, large batches of steady-state chunks (Tq <= 8 for every entry): FOUR adjacent channels per thread, one CTA per entry.
// dwconv2_kernel reads its 18 filter taps as scalars 72 bytes apart across the lanes of a warp (11.3 M load sectors per launch at
// 1024 streams) and slides its 9-tap window through register moves inside a loop over the rows: 10 M warp instructions, 22 us
// (profiles/r02_misc...).  Here the taps come from a transposed copy [9][1024] as float4 (one 512-byte run per warp and tap,
// requested before the dependency wait: constants), the extended signal [cache(4) | rows | 0000] of a channel pair lives in registers
// at fixed positions (rows fully unrolled, no window shifting), the multiply-adds run on the packed f32x2 pipe (FFMA2: channel pairs),
// activations move as 8-byte bf16 quads and the time cache as 64 contiguous bytes per thread.  Per channel the arithmetic and its
// order are exactly those of dwconv_kernel.
__global__ void __launch_bounds__(256)
dwconv4_kernel(BatchDev b, DwConvArgs a) {
  pdl_trigger();
  const int e = blockIdx.x, ch = 4 * threadIdx.x;
  float2 w01[kConvK], w23[kConvK];
#pragma unroll
  for (int i = 0; i < kConvK; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(a.wt + (size_t)i * kDModel + ch);
    w01[i] = make_float2(t.x, t.y); w23[i] = make_float2(t.z, t.w);
  }
  const float4 bias4 = *reinterpret_cast<const float4*>(a.bias + ch);
  pdl_wait();
  const int Tq = b.Tq[e], qlen = b.qlen[e], row0 = b.row_off[e];
  const int nv = Tq < qlen ? Tq : qlen;                   // rows past qlen are padding: zero (pad_mask)
  float* cache = a.cache_tm + (size_t)b.slot[e] * a.slot_stride + (size_t)ch * kTimeCtx;
  const bool offline = b.offline[e] != 0;
  // ext[0..16) = [cache(4) | c rows 0..7 | 0 0 0 0] for the channel pairs (ch, ch+1) and (ch+2, ch+3)
  float2 e01[16], e23[16];
  {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 c0 = offline ? z4 : *reinterpret_cast<const float4*>(cache);
    const float4 c1 = offline ? z4 : *reinterpret_cast<const float4*>(cache + kTimeCtx);
    const float4 c2 = offline ? z4 : *reinterpret_cast<const float4*>(cache + 2 * kTimeCtx);
    const float4 c3 = offline ? z4 : *reinterpret_cast<const float4*>(cache + 3 * kTimeCtx);
    e01[0] = make_float2(c0.x, c1.x); e01[1] = make_float2(c0.y, c1.y); e01[2] = make_float2(c0.z, c1.z); e01[3] = make_float2(c0.w, c1.w);
    e23[0] = make_float2(c2.x, c3.x); e23[1] = make_float2(c2.y, c3.y); e23[2] = make_float2(c2.z, c3.z); e23[3] = make_float2(c2.w, c3.w);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    uint2 raw = make_uint2(0u, 0u);
    if (t < nv) raw = *reinterpret_cast<const uint2*>(a.c_bf16 + (size_t)(row0 + t) * kDModel + ch);
    e01[4 + t] = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
    e23[4 + t] = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
  }
#pragma unroll
  for (int i = 12; i < 16; ++i) { e01[i] = make_float2(0.f, 0.f); e23[i] = make_float2(0.f, 0.f); }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (t < Tq) {
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kConvK; ++i) { a01 = __ffma2_rn(w01[i], e01[t + i], a01); a23 = __ffma2_rn(w23[i], e23[t + i], a23); }
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(silu(a01.x + bias4.x), silu(a01.y + bias4.y));
      const __nv_bfloat162 h1 = __floats2bfloat162_rn(silu(a23.x + bias4.z), silu(a23.y + bias4.w));
      *reinterpret_cast<uint2*>(a.out.ptr + (size_t)(row0 + t) * a.out.lda + ch) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    }
  }
  if (!offline) {
    // new cache = ext[Tq+1 .. Tq+5)   (new_x[:-3][-4:], cache_drop_size 3); Tq is uniform over the CTA
    float2 n01[4], n23[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { n01[i] = e01[1 + i]; n23[i] = e23[1 + i]; }
#pragma unroll
    for (int tt = 1; tt <= 8; ++tt) {
      if (Tq == tt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { n01[i] = e01[tt + 1 + i]; n23[i] = e23[tt + 1 + i]; }
      }
    }
    *reinterpret_cast<float4*>(cache) = make_float4(n01[0].x, n01[1].x, n01[2].x, n01[3].x);
    *reinterpret_cast<float4*>(cache + kTimeCtx) = make_float4(n01[0].y, n01[1].y, n01[2].y, n01[3].y);
    *reinterpret_cast<float4*>(cache + 2 * kTimeCtx) = make_float4(n23[0].x, n23[1].x, n23[2].x, n23[3].x);
    *reinterpret_cast<float4*>(cache + 3 * kTimeCtx) = make_float4(n23[0].y, n23[1].y, n23[2].y, n23[3].y);
  }
}
void launch_dwconv(const BatchDev& b, const DwConvArgs& a, cudaStream_t st) {
  static const bool quad = [] { const char* v = getenv("PARAKEET_B200_DWCONV4"); return !(v && v[0] == '0'); }();
  if (quad && a.wt != nullptr && a.c_bf16 != nullptr && a.out.lo_off == 0 && b.B >= 64 && b.max_Tq <= 8) {
    launch_k(dwconv4_kernel, dim3(b.B), dim3(256), 0, st, b, a);
    PKB_CUDA(cudaGetLastError());
    return;
  }
  if (b.B <= 0) return;
  static const bool pair = [] { const char* v = getenv("PARAKEET_B200_DWCONV2"); return !(v && v[0] == '0'); }();
  if (pair && a.c_bf16 != nullptr && a.out.lo_off == 0) launch_k(dwconv2_kernel, dim3(b.B, kDModel / 512), dim3(256), 0, st, b, a);
  else launch_k(dwconv_kernel, dim3(b.B, kDModel / 256), dim3(256), 0, st, b, a);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ output
__global__ void __launch_bounds__(256)
gather_output_kernel(BatchDev b, const float* __restrict__ x, float* __restrict__ enc_out, int out_T) {
  pdl_enter();
  const int e = blockIdx.x;
  const int row0 = b.row_off[e], Tq = b.Tq[e];
  for (int i = threadIdx.x; i < kDModel * out_T; i += 256) {
    const int d = i / out_T, t = i % out_T;
    enc_out[(size_t)e * kDModel * out_T + i] = t < Tq ? x[(size_t)(row0 + t) * kDModel + d] : 0.0f;
  }
}
void launch_gather_output(const BatchDev& b, const float* x, float* enc_out, int out_T, cudaStream_t st) {
  if (b.B <= 0) return;
  launch_k(gather_output_kernel, dim3(b.B), dim3(256), 0, st, b, x, enc_out, out_T);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ import / export
__global__ void __launch_bounds__(256)
acache_to_act_kernel(const void* __restrict__ acache, int is_f32, const int* __restrict__ slots, const int* __restrict__ heads,
                     ActOut a) {
  const int e = blockIdx.y, j = blockIdx.x;     // logical cache row j of entry e
  const int phys = (heads[e] + j) % kRingCap;
  const size_t src = ((size_t)slots[e] * kRingCap + phys) * kDModel;
  for (int c = threadIdx.x; c < kDModel; c += 256) {
    const float v = is_f32 ? ((const float*)acache)[src + c] : __bfloat162float(((const __nv_bfloat16*)acache)[src + c]);
    store_act(a.ptr, (size_t)e * kCacheS + j, a.lda, c, v, a.lo_off);
  }
}
void launch_acache_to_act(const void* acache_layer, int is_f32, const int* slots, const int* heads, int n, ActOut a, cudaStream_t st) {
  if (n <= 0) return;
  acache_to_act_kernel<<<dim3(kCacheS, n), 256, 0, st>>>(acache_layer, is_f32, slots, heads, a);
  PKB_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256)
acache_copy_kernel(void* __restrict__ acache, int is_f32, const int* __restrict__ slots, const int* __restrict__ heads,
                   float* __restrict__ ext, long long ext_entry_stride, int to_ring) {
  const int e = blockIdx.y, j = blockIdx.x;
  const int phys = (heads[e] + j) % kRingCap;
  const size_t r = ((size_t)slots[e] * kRingCap + phys) * kDModel;
  float* x = ext + (size_t)e * ext_entry_stride + (size_t)j * kDModel;
  for (int c = threadIdx.x; c < kDModel; c += 256) {
    if (to_ring) {
      if (is_f32) ((float*)acache)[r + c] = x[c];
      else ((__nv_bfloat16*)acache)[r + c] = __float2bfloat16_rn(x[c]);
    } else {
      x[c] = is_f32 ? ((const float*)acache)[r + c] : __bfloat162float(((const __nv_bfloat16*)acache)[r + c]);
    }
  }
}
void launch_acache_import(void* acache_layer, int is_f32, const int* slots, const int* heads, int n, const float* src,
                          long long src_entry_stride, cudaStream_t st) {
  if (n <= 0) return;
  acache_copy_kernel<<<dim3(kCacheS, n), 256, 0, st>>>(acache_layer, is_f32, slots, heads, const_cast<float*>(src),
                                                        src_entry_stride, 1);
  PKB_CUDA(cudaGetLastError());
}
void launch_acache_export(const void* acache_layer, int is_f32, const int* slots, const int* heads, int n, float* dst,
                          long long dst_entry_stride, cudaStream_t st) {
  if (n <= 0) return;
  acache_copy_kernel<<<dim3(kCacheS, n), 256, 0, st>>>(const_cast<void*>(acache_layer), is_f32, slots, heads, dst,
                                                        dst_entry_stride, 0);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
