// offline_long.cu -- see offline_long.cuh.
//
// Semantics restated from NeMo 2.6.0 (ConformerEncoder.forward with full context, the graph the reference exports as
// `encoder`, /root/reference/tools/export_onnx/export.py:343-375 and contract.json:67-96):
//   * RelPositionalEncoding: 2T-1 rows for relative positions T-1 ... -(T-1)
//   * RelPositionMultiHeadAttention: scores[i][j] = ((q_i+u).k_j + (q_i+v).p_{i-j}) / sqrt(d_k)   (rel_shift == position i-j)
//   * ConformerConvolution with symmetric (4,4) zero padding
// B200-first layout: the position term is not materialised as a [T, 2T-1] matrix (64 GB per layer for a 1 h clip); each
// 64 x 64 score tile multiplies the 64 queries with the 127-row window of the projected table that the tile can touch and
// applies the rel_shift skew by index inside the CTA.
#include <cuda.h>

#include "offline_long.cuh"

namespace pkb {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

constexpr int kBM = 64;            // query rows per CTA (16 per warp)
constexpr int kBN = 64;            // keys per tile
constexpr int kPitch = 136;        // bf16 elements per staged row (272 B = 17 x 16 B: conflict-free ldmatrix)
constexpr int kWinRows = 128;      // position-table window rows staged per tile (127 used)
constexpr int kGPitchF = 88;       // floats per row of a warp's skew buffer (80 window columns; 88 % 32 == 24: conflict-free float2 stores)
constexpr int kGCols = 80;         // window columns one warp's 16 rows can touch (16 + 64 - 1 = 79, rounded to n-tiles of 8)
constexpr size_t kLfSmem = (size_t)(2 * kBN + kWinRows) * kPitch * 2 + (size_t)4 * 16 * kGPitchF * 4;
constexpr float kScale = 0.08838834764831845f;      // 1/sqrt(128)

// div_i table (512 floats), filled once from the host so that it carries the host libm's expf (as the streaming table does)
__constant__ float c_lf_div[kDModel / 2];

}  // namespace

// ------------------------------------------------------------------------------------------------ position table
__global__ void __launch_bounds__(256)
lf_posemb_kernel(ActOut a, int Tm) {
  pdl_enter();
  const int r = blockIdx.x;
  const float pos = (float)(r - (Tm - 1));
  for (int i = threadIdx.x; i < kDModel / 2; i += 256) {
    const float ang = pos * c_lf_div[i];
    store_act(a.ptr, r, a.lda, 2 * i, sinf(ang), a.lo_off);
    store_act(a.ptr, r, a.lda, 2 * i + 1, cosf(ang), a.lo_off);
  }
}
void launch_lf_posemb(ActOut a, int Tm, cudaStream_t st) {
  static bool init = false;
  if (!init) {
    float div[kDModel / 2];
    for (int i = 0; i < kDModel / 2; ++i) div[i] = expf((float)(2 * i) * -(logf(10000.0f) / (float)kDModel));
    PKB_CUDA(cudaMemcpyToSymbol(c_lf_div, div, sizeof(div)));
    init = true;
  }
  if (Tm <= 0) return;
  launch_k(lf_posemb_kernel, dim3(2 * Tm - 1), dim3(256), 0, st, a, Tm);
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ attention, bf16 tensor cores
// grid (query tile, head, entry); 4 warps, warp w owns query rows [16w, 16w+16) of the tile.
// Per key tile:  S = Qu K^T (64 MMAs per warp),  G = Qv Pwin^T over the warp's 80 window columns (80 MMAs), skew
// S[r][c] += G[r][r - c + 63] through a per-warp smem buffer, online softmax in registers, O += P V (64 MMAs).
__global__ void __launch_bounds__(128, 2)
lf_attention_mma_kernel(BatchDev b, LfAttnArgs a) {
  extern __shared__ __align__(16) unsigned char lf_smem[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(lf_smem);                 // [64][kPitch]
  __nv_bfloat16* sV = sK + kBN * kPitch;                                          // [64][kPitch]
  __nv_bfloat16* sP = sV + kBN * kPitch;                                          // [128][kPitch]; first the Qu / Qv staging
  float* sG = reinterpret_cast<float*>(sP + kWinRows * kPitch);                   // [4 warps][16][kGPitchF]
  pdl_enter();
  const int e = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kBM;
  const int T = b.Tq[e];
  if (i0 >= T) return;
  const int row0 = b.row_off[e];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* qkv = a.qkv_bf16 + (size_t)row0 * (3 * kDModel) + h * kDHead;

  // ---- stage Qu = q + pos_bias_u and Qv = q + pos_bias_v (bf16) in the window buffer, pull them into A fragments
  for (int x = tid; x < kBM * 16; x += 128) {
    const int r = x >> 4, c8 = (x & 15) * 8;
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    const bool ok = i0 + r < T;
    if (ok) raw = *reinterpret_cast<const uint4*>(qkv + (size_t)(i0 + r) * (3 * kDModel) + c8);
    const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
    uint32_t ou[4], ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 q = __bfloat1622float2(q2[k]);
      const int d = h * kDHead + c8 + 2 * k;
      ou[k] = ok ? pack_bf16x2(q.x + a.bias_u[d], q.y + a.bias_u[d + 1]) : 0u;
      ov[k] = ok ? pack_bf16x2(q.x + a.bias_v[d], q.y + a.bias_v[d + 1]) : 0u;
    }
    *reinterpret_cast<uint4*>(sP + r * kPitch + c8) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
    *reinterpret_cast<uint4*>(sP + (kBM + r) * kPitch + c8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
  }
  __syncthreads();
  uint32_t qu[8][4], qv[8][4];
  {
    const int r = 16 * warp + (lane & 15), c = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      ldsm_x4(smem_u32(sP + r * kPitch + 16 * ks + c), qu[ks][0], qu[ks][1], qu[ks][2], qu[ks][3]);
      ldsm_x4(smem_u32(sP + (kBM + r) * kPitch + 16 * ks + c), qv[ks][0], qv[ks][1], qv[ks][2], qv[ks][3]);
    }
  }
  __syncthreads();

  float o[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float* sGw = sG + warp * 16 * kGPitchF;
  const int bl_row = lane & 7, bl_chunk = lane >> 3;       // ldmatrix lane -> (row, 16-byte chunk) for B operands
  const int n_ppos_rows = 2 * a.Tm - 1;

#pragma unroll 1
  for (int j0 = 0; j0 < T; j0 += kBN) {
    // ---- stage K, V (rows past the utterance: zero) and the 127-row table window
    for (int x = tid; x < kBN * 16; x += 128) {
      const int r = x >> 4, c8 = (x & 15) * 8;
      uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
      if (j0 + r < T) {
        const __nv_bfloat16* src = qkv + (size_t)(j0 + r) * (3 * kDModel) + c8;
        kk = *reinterpret_cast<const uint4*>(src + kDModel);
        vv = *reinterpret_cast<const uint4*>(src + 2 * kDModel);
      }
      *reinterpret_cast<uint4*>(sK + r * kPitch + c8) = kk;
      *reinterpret_cast<uint4*>(sV + r * kPitch + c8) = vv;
    }
    const int rel0 = i0 - j0 - (kBN - 1);                  // relative position of window row 0
    for (int x = tid; x < kWinRows * 16; x += 128) {
      const int r = x >> 4, c8 = (x & 15) * 8;
      const int prow = rel0 + r + (a.Tm - 1);
      uint4 pv = make_uint4(0u, 0u, 0u, 0u);
      if (prow >= 0 && prow < n_ppos_rows) pv = *reinterpret_cast<const uint4*>(a.ppos_bf16 + (size_t)prow * kDModel + h * kDHead + c8);
      *reinterpret_cast<uint4*>(sP + r * kPitch + c8) = pv;
    }
    __syncthreads();

    // ---- position scores of this warp's rows over its 80 window columns -> skew buffer
#pragma unroll 1
    for (int nt = 0; nt < kGCols / 8; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const __nv_bfloat16* prow = sP + (16 * warp + 8 * nt + bl_row) * kPitch + 8 * bl_chunk;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(smem_u32(prow + 32 * kp), r0, r1, r2, r3);
        mma_bf16(c, qv[2 * kp], r0, r1);
        mma_bf16(c, qv[2 * kp + 1], r2, r3);
      }
      *reinterpret_cast<float2*>(sGw + g * kGPitchF + 8 * nt + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(sGw + (g + 8) * kGPitchF + 8 * nt + 2 * t) = make_float2(c[2], c[3]);
    }
    __syncwarp();

    // ---- content scores + skewed position term, scale, mask
    float s[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
      const __nv_bfloat16* krow = sK + (8 * nb + bl_row) * kPitch + 8 * bl_chunk;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(smem_u32(krow + 32 * kp), r0, r1, r2, r3);
        mma_bf16(s[nb], qu[2 * kp], r0, r1);
        mma_bf16(s[nb], qu[2 * kp + 1], r2, r3);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int rl = g + 8 * (x >> 1), c = 8 * nb + 2 * t + (x & 1);       // row inside the warp's 16, key inside the tile
        const float gpos = sGw[rl * kGPitchF + (rl - c + (kBN - 1))];
        const float v = (j0 + c < T) ? (s[nb][x] + gpos) * kScale : -INFINITY;
        s[nb][x] = v;
        mx[x >> 1] = fmaxf(mx[x >> 1], v);
      }
    // ---- online softmax (rows g and g+8 of the warp; the 4 lanes of a quad share a row)
    float alpha[2];
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      mx[hr] = fmaxf(mx[hr], __shfl_xor_sync(0xffffffffu, mx[hr], 1));
      mx[hr] = fmaxf(mx[hr], __shfl_xor_sync(0xffffffffu, mx[hr], 2));
      const float m_new = fmaxf(m_run[hr], mx[hr]);          // finite from the first tile on (key 0 is always valid)
      alpha[hr] = __expf(m_run[hr] - m_new);                 // first tile: exp(-inf) = 0
      m_run[hr] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pa[4][4];                                       // probabilities as A fragments: k-step kk covers keys 16kk .. 16kk+15
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = __expf(s[nb][0] - m_run[0]), p1 = __expf(s[nb][1] - m_run[0]);
      const float p2 = __expf(s[nb][2] - m_run[1]), p3 = __expf(s[nb][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      rs[hr] += __shfl_xor_sync(0xffffffffu, rs[hr], 1);
      rs[hr] += __shfl_xor_sync(0xffffffffu, rs[hr], 2);
      l_run[hr] = l_run[hr] * alpha[hr] + rs[hr];
    }
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
      o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const __nv_bfloat16* vrow = sV + (16 * kk + (lane & 15)) * kPitch + (lane >> 4) * 8;
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(smem_u32(vrow + 16 * np), r0, r1, r2, r3);
        mma_bf16(o[2 * np], pa[kk], r0, r1);
        mma_bf16(o[2 * np + 1], pa[kk], r2, r3);
      }
    }
    __syncthreads();      // every warp is done with K, V and the window before the next tile overwrites them
  }

  // ---- context rows -> bf16 operand of linear_out
#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    const int i = i0 + 16 * warp + g + 8 * hr;
    if (i >= T) continue;
    const float inv = l_run[hr] > 0.f ? 1.f / l_run[hr] : 0.f;
    __nv_bfloat16* dst = a.ctx.ptr + (size_t)(row0 + i) * a.ctx.lda + h * kDHead + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
      *reinterpret_cast<uint32_t*>(dst + 8 * nt) = pack_bf16x2(o[nt][2 * hr] * inv, o[nt][2 * hr + 1] * inv);
  }
}

// ------------------------------------------------------------------------------------------------ attention, TMA-staged version
// Same tiling and arithmetic as lf_attention_mma_kernel, but the K, V and table-window tiles arrive by TMA (64-dim x 64/128-row
// boxes, 128-byte swizzle -> conflict-free ldmatrix without padding) instead of LDG + STS through the LSU: the first version
// was bound by the shared-memory data pipe (l1tex lsu wavefronts 67 % of peak, tensor pipe 37 %; profiles/r01_lfattn_r1k), and a
// third of those wavefronts were the staging copies.  Single buffers, split in two phases so the copies still overlap the math:
// K and the window are free again after the score phase (their next tile loads during softmax + PV), V after the PV phase (its
// next tile loads during the next score phase).
namespace {
constexpr int kBoxK = kBN * 128;                 // one 64-row x 64-dim bf16 box = 8 KB
constexpr int kBoxP = kWinRows * 128;            // one 128-row x 64-dim box = 16 KB
constexpr int kOffV = 2 * kBoxK, kOffP = 4 * kBoxK, kOffG = kOffP + 2 * kBoxP, kOffBar = kOffG + 4 * 16 * kGPitchF * 4;
constexpr size_t kLfSmemTma = 1024 + kOffBar + 64;
__device__ __forceinline__ uint32_t swz(uint32_t tile, int box_bytes, int row, int chunk16) {      // (row, 16-byte chunk 0..15) inside a two-box tile
  return tile + (chunk16 >> 3) * box_bytes + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4);
}
}  // namespace

__global__ void __launch_bounds__(128, 2)
lf_attention_tma_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_pos, BatchDev b, LfAttnArgs a) {
  extern __shared__ __align__(16) unsigned char lf_smem[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)lf_smem + 1023) & ~(uintptr_t)1023);
  float* sG = reinterpret_cast<float*>(base + kOffG);
  uint64_t* bar_kp = reinterpret_cast<uint64_t*>(base + kOffBar);
  uint64_t* bar_v = bar_kp + 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qkv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_pos) : "memory");
    mbar_init(bar_kp, 1);
    mbar_init(bar_v, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_enter();
  const int e = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kBM;
  const int T = b.Tq[e];
  if (i0 >= T) return;
  const int row0 = b.row_off[e];
  const __nv_bfloat16* qkv = a.qkv_bf16 + (size_t)row0 * (3 * kDModel) + h * kDHead;

  // ---- stage Qu / Qv (bf16, 256-byte rows) in the window buffer, pull them into A fragments
  {
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(base + kOffP);
    for (int x = tid; x < kBM * 16; x += 128) {
      const int r = x >> 4, c8 = (x & 15) * 8;
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      const bool ok = i0 + r < T;
      if (ok) raw = *reinterpret_cast<const uint4*>(qkv + (size_t)(i0 + r) * (3 * kDModel) + c8);
      const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
      uint32_t ou[4], ov[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 q = __bfloat1622float2(q2[k]);
        const int d = h * kDHead + c8 + 2 * k;
        ou[k] = ok ? pack_bf16x2(q.x + a.bias_u[d], q.y + a.bias_u[d + 1]) : 0u;
        ov[k] = ok ? pack_bf16x2(q.x + a.bias_v[d], q.y + a.bias_v[d + 1]) : 0u;
      }
      *reinterpret_cast<uint4*>(sQ + r * kDHead + c8) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<uint4*>(sQ + (kBM + r) * kDHead + c8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    }
  }
  __syncthreads();
  uint32_t qu[8][4], qv[8][4];
  {
    const __nv_bfloat16* sQ = reinterpret_cast<const __nv_bfloat16*>(base + kOffP);
    const int r = 16 * warp + (lane & 15), c = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      ldsm_x4(smem_u32(sQ + r * kDHead + 16 * ks + c), qu[ks][0], qu[ks][1], qu[ks][2], qu[ks][3]);
      ldsm_x4(smem_u32(sQ + (kBM + r) * kDHead + 16 * ks + c), qv[ks][0], qv[ks][1], qv[ks][2], qv[ks][3]);
    }
  }
  __syncthreads();

  const int n_tiles = (T + kBN - 1) / kBN;
  const int col_k = kDModel + h * kDHead, col_v = 2 * kDModel + h * kDHead, col_p = h * kDHead;
  auto issue_kp = [&](int tile) {
    const int j0 = tile * kBN;
    mbar_expect_tx(bar_kp, 2 * kBoxK + 2 * kBoxP);
    tma_load_2d(base, &map_qkv, bar_kp, col_k, row0 + j0);
    tma_load_2d(base + kBoxK, &map_qkv, bar_kp, col_k + 64, row0 + j0);
    const int prow = i0 - j0 - (kBN - 1) + (a.Tm - 1);        // table row of window row 0 (rows outside the table read as zero)
    tma_load_2d(base + kOffP, &map_pos, bar_kp, col_p, prow);
    tma_load_2d(base + kOffP + kBoxP, &map_pos, bar_kp, col_p + 64, prow);
  };
  auto issue_v = [&](int tile) {
    mbar_expect_tx(bar_v, 2 * kBoxK);
    tma_load_2d(base + kOffV, &map_qkv, bar_v, col_v, row0 + tile * kBN);
    tma_load_2d(base + kOffV + kBoxK, &map_qkv, bar_v, col_v + 64, row0 + tile * kBN);
  };
  if (tid == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the Q staging (generic proxy) precedes TMA writes to the same bytes
    issue_kp(0);
    issue_v(0);
  }

  float o[16][4];
#pragma unroll
  for (int nt = 0; nt < 16; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float* sGw = sG + warp * 16 * kGPitchF;
  const int bl_row = lane & 7, bl_chunk = lane >> 3;
  const uint32_t uK = smem_u32(base), uV = smem_u32(base + kOffV), uP = smem_u32(base + kOffP);

#pragma unroll 1
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int j0 = tile * kBN;
    const uint32_t ph = tile & 1;
    mbar_wait(bar_kp, ph);
    // ---- position scores of this warp's rows over its 80 window columns -> skew buffer
#pragma unroll 1
    for (int nt = 0; nt < kGCols / 8; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const int prow = 16 * warp + 8 * nt + bl_row;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(swz(uP, kBoxP, prow, 4 * kp + bl_chunk), r0, r1, r2, r3);
        mma_bf16(c, qv[2 * kp], r0, r1);
        mma_bf16(c, qv[2 * kp + 1], r2, r3);
      }
      *reinterpret_cast<float2*>(sGw + g * kGPitchF + 8 * nt + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(sGw + (g + 8) * kGPitchF + 8 * nt + 2 * t) = make_float2(c[2], c[3]);
    }
    __syncwarp();
    // ---- content scores
    float s[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
      const int krow = 8 * nb + bl_row;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(swz(uK, kBoxK, krow, 4 * kp + bl_chunk), r0, r1, r2, r3);
        mma_bf16(s[nb], qu[2 * kp], r0, r1);
        mma_bf16(s[nb], qu[2 * kp + 1], r2, r3);
      }
    }
    __syncthreads();                 // every warp is done with K and the window
    if (tid == 0 && tile + 1 < n_tiles) issue_kp(tile + 1);
    // ---- skewed position term, scale, mask
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int rl = g + 8 * (x >> 1), c = 8 * nb + 2 * t + (x & 1);
        const float gpos = sGw[rl * kGPitchF + (rl - c + (kBN - 1))];
        const float v = (j0 + c < T) ? (s[nb][x] + gpos) * kScale : -INFINITY;
        s[nb][x] = v;
        mx[x >> 1] = fmaxf(mx[x >> 1], v);
      }
    float alpha[2];
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      mx[hr] = fmaxf(mx[hr], __shfl_xor_sync(0xffffffffu, mx[hr], 1));
      mx[hr] = fmaxf(mx[hr], __shfl_xor_sync(0xffffffffu, mx[hr], 2));
      const float m_new = fmaxf(m_run[hr], mx[hr]);
      alpha[hr] = __expf(m_run[hr] - m_new);
      m_run[hr] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pa[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = __expf(s[nb][0] - m_run[0]), p1 = __expf(s[nb][1] - m_run[0]);
      const float p2 = __expf(s[nb][2] - m_run[1]), p3 = __expf(s[nb][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      rs[hr] += __shfl_xor_sync(0xffffffffu, rs[hr], 1);
      rs[hr] += __shfl_xor_sync(0xffffffffu, rs[hr], 2);
      l_run[hr] = l_run[hr] * alpha[hr] + rs[hr];
    }
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
      o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
    }
    // ---- O += P V
    mbar_wait(bar_v, ph);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int vrow = 16 * kk + (lane & 15);
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(swz(uV, kBoxK, vrow, 2 * np + (lane >> 4)), r0, r1, r2, r3);
        mma_bf16(o[2 * np], pa[kk], r0, r1);
        mma_bf16(o[2 * np + 1], pa[kk], r2, r3);
      }
    }
    __syncthreads();                 // every warp is done with V
    if (tid == 0 && tile + 1 < n_tiles) issue_v(tile + 1);
  }

#pragma unroll
  for (int hr = 0; hr < 2; ++hr) {
    const int i = i0 + 16 * warp + g + 8 * hr;
    if (i >= T) continue;
    const float inv = l_run[hr] > 0.f ? 1.f / l_run[hr] : 0.f;
    __nv_bfloat16* dst = a.ctx.ptr + (size_t)(row0 + i) * a.ctx.lda + h * kDHead + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
      *reinterpret_cast<uint32_t*>(dst + 8 * nt) = pack_bf16x2(o[nt][2 * hr] * inv, o[nt][2 * hr + 1] * inv);
  }
}

// ------------------------------------------------------------------------------------------------ attention, f32 CUDA cores
// One CTA (128 threads) per (query row, head, entry); the row's T scores live in shared memory.
__global__ void __launch_bounds__(128)
lf_attention_f32_kernel(BatchDev b, LfAttnArgs a) {
  extern __shared__ __align__(16) unsigned char lf_smem[];
  float* s_sc = reinterpret_cast<float*>(lf_smem);       // [T]
  __shared__ __align__(16) float s_qu[kDHead], s_qv[kDHead];
  __shared__ float s_red[4];
  pdl_enter();
  const int e = blockIdx.z, h = blockIdx.y, i = blockIdx.x;
  const int T = b.Tq[e];
  if (i >= T) return;
  const int row0 = b.row_off[e];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qkv = a.qkv_f32 + (size_t)row0 * (3 * kDModel) + h * kDHead;
  {
    const float q = qkv[(size_t)i * (3 * kDModel) + tid];
    s_qu[tid] = q + a.bias_u[h * kDHead + tid];
    s_qv[tid] = q + a.bias_v[h * kDHead + tid];
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < T; j += 128) {
    const float4* k4 = reinterpret_cast<const float4*>(qkv + (size_t)j * (3 * kDModel) + kDModel);
    const float4* p4 = reinterpret_cast<const float4*>(a.ppos_f32 + (size_t)(i - j + a.Tm - 1) * kDModel + h * kDHead);
    float ac = 0.f, bd = 0.f;
#pragma unroll 8
    for (int d4 = 0; d4 < kDHead / 4; ++d4) {
      const float4 kv = k4[d4], pv = p4[d4];
      const float4 u = *reinterpret_cast<const float4*>(s_qu + 4 * d4), v = *reinterpret_cast<const float4*>(s_qv + 4 * d4);
      ac = fmaf(u.x, kv.x, ac); ac = fmaf(u.y, kv.y, ac); ac = fmaf(u.z, kv.z, ac); ac = fmaf(u.w, kv.w, ac);
      bd = fmaf(v.x, pv.x, bd); bd = fmaf(v.y, pv.y, bd); bd = fmaf(v.z, pv.z, bd); bd = fmaf(v.w, pv.w, bd);
    }
    const float sc = (ac + bd) * kScale;
    s_sc[j] = sc;
    mx = fmaxf(mx, sc);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < T; j += 128) {
    const float p = expf(s_sc[j] - mx);
    s_sc[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  // thread == head dim
  const float* vcol = qkv + 2 * kDModel + tid;
  float acc0 = 0.f, acc1 = 0.f;
  int j = 0;
  for (; j + 1 < T; j += 2) {
    acc0 = fmaf(s_sc[j], vcol[(size_t)j * (3 * kDModel)], acc0);
    acc1 = fmaf(s_sc[j + 1], vcol[(size_t)(j + 1) * (3 * kDModel)], acc1);
  }
  if (j < T) acc0 = fmaf(s_sc[j], vcol[(size_t)j * (3 * kDModel)], acc0);
  store_act(a.ctx.ptr, row0 + i, a.ctx.lda, h * kDHead + tid, (acc0 + acc1) * inv, a.ctx.lo_off);
}

void launch_lf_attention(const BatchDev& b, const LfAttnArgs& a, cudaStream_t st) {
  if (b.B <= 0 || a.max_T <= 0) return;
  if (a.qkv_bf16) {
    PKB_CHECK(a.ctx.lo_off == 0 && a.ppos_bf16, "lf_attention: the tensor-core kernel is the bf16-mode path");
    static bool attr = false;
    if (!attr) {
      PKB_CUDA(cudaFuncSetAttribute(lf_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLfSmem));
      PKB_CUDA(cudaFuncSetAttribute(lf_attention_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLfSmemTma));
      attr = true;
    }
    static const bool use_tma = [] { const char* v = getenv("PARAKEET_B200_LF_ATTN_TMA"); return !(v && v[0] == '0'); }();
    const dim3 grid((a.max_T + kBM - 1) / kBM, kHeads, b.B);
    if (use_tma && a.map_qkv && a.map_pos)
      launch_k(lf_attention_tma_kernel, grid, dim3(128), kLfSmemTma, st, *reinterpret_cast<const CUtensorMap*>(a.map_qkv),
               *reinterpret_cast<const CUtensorMap*>(a.map_pos), b, a);
    else
      launch_k(lf_attention_mma_kernel, grid, dim3(128), kLfSmem, st, b, a);
  } else {
    PKB_CHECK(a.qkv_f32 && a.ppos_f32, "lf_attention: precise mode needs f32 q/k/v and table");
    const size_t smem = (size_t)a.max_T * sizeof(float);
    PKB_CHECK(smem <= 200 * 1024, "lf_attention (precise mode): utterance too long for the f32 kernel (max 51200 encoder frames)");
    static size_t attr_bytes = 0;
    if (smem > attr_bytes) {
      PKB_CUDA(cudaFuncSetAttribute(lf_attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
      attr_bytes = 200 * 1024;
    }
    launch_k(lf_attention_f32_kernel, dim3(a.max_T, kHeads, b.B), dim3(128), smem, st, b, a);
  }
  PKB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ conv module middle
__global__ void __launch_bounds__(256)
lf_dwconv_kernel(BatchDev b, LfDwConvArgs a) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int m = (int)(idx >> 8), c = (int)(idx & 255) * 4;
  if (m >= a.M) return;
  const int e = b.row_entry[m], t = b.row_pos[m], T = b.Tq[e];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < kConvK; ++k) {
    const int tt = t + k - 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (tt >= 0 && tt < T) {
      const size_t i = (size_t)(m + k - 4) * kDModel + c;
      if (a.c_bf16) {
        const uint2 raw = *reinterpret_cast<const uint2*>(a.c_bf16 + i);
        const float2 p0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
        const float2 p1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
        v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y;
      } else {
        const float4 f = *reinterpret_cast<const float4*>(a.c + i);
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) acc[x] = fmaf(a.w[(c + x) * kConvK + k], v[x], acc[x]);
  }
  const float4 bs = *reinterpret_cast<const float4*>(a.bias + c);
  store_act4(a.out.ptr, m, a.out.lda, c, make_float4(silu(acc[0] + bs.x), silu(acc[1] + bs.y), silu(acc[2] + bs.z), silu(acc[3] + bs.w)),
             a.out.lo_off);
}
void launch_lf_dwconv(const BatchDev& b, const LfDwConvArgs& a, cudaStream_t st) {
  if (a.M <= 0) return;
  const long long n = (long long)a.M * 256;
  launch_k(lf_dwconv_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, b, a);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
