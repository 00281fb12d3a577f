// gemm_tc.cu -- tcgen05 / TMA / TMEM GEMM for sm_100a (see gemm.h):  C[m,n] = sum_k A[m,k] * W[n,k], fused epilogue.
//
// Persistent, warp-specialised kernel: one CTA per SM loops over 128 x BN output tiles (BN = 128 or 256, m fastest so
// that concurrently running CTAs share weight tiles through L2).  320 threads:
//   warp 0    : TMA producer  -- per k-block one 128x64 bf16 A tile and BN/128 128x64 W tiles (cp.async.bulk.tensor.2d,
//               128-byte swizzle) into a kStages-deep shared-memory ring; completion on "full" mbarriers
//   warp 1    : MMA issuer    -- one lane issues 4 x tcgen05.mma (M128 N{128,256} K16, bf16 -> f32) per k-block; the
//               accumulator lives in TMEM and is DOUBLE-BUFFERED (2 x BN columns): tile i+1 is accumulated while the
//               epilogue drains tile i.  tcgen05.commit releases smem slots ("empty") and publishes tiles ("tmem_full").
//   warps 2-9 : epilogue      -- warp w reads TMEM lane quarter w%4, column half (w-2)/4: tcgen05.ld 32 lanes x 16 columns,
//               transposes the 32x16 block through a private padded smem tile so that every global store instruction
//               writes whole 64-byte row segments (8 rows per instruction), applies the fused epilogue (bias / ReLU /
//               SiLU / GLU / residual add / row map / QKV scatter into the per-stream rings) on float4 granules, and
//               hands the TMEM buffer back ("tmem_empty") as soon as its last tcgen05.ld has landed.
// Split ("precise") mode runs the k-loop twice over the same W tiles: first the bf16 high plane of A, then the low
// plane, accumulating into the same TMEM tile -- fp32-grade products at 2x the tensor work, no extra weight traffic
// from HBM (the W tile of the second pass hits L2).
#include <cuda.h>

#include "gemm.h"

namespace pkb {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kTileABytes = BM * BK * 2;          // 16 KB
constexpr int kEpiPitch = 20;                     // floats per staged row (16 + 4 pad: conflict-free 128-bit access)
constexpr int kEpiBytesPerWarp = 32 * kEpiPitch * 4;

template <int BN>
struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kTileBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kTileABytes + kTileBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr size_t kSmemBytes = 1024 + (size_t)kStages * kStageBytes + (size_t)kEpiWarps * kEpiBytesPerWarp + 256;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: 8-row groups are 1024 B apart (SBO), LBO unused, descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, 16-byte units
  d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset
  d |= (uint64_t)1 << 46;                               // version
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
template <int BN>
struct IDesc {
  static constexpr uint32_t value = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc_v, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc_v), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"   // same asm statement: the registers are only defined once the load has landed
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// One 256-bit store (STG.E.256, sm_100): a lane writes a whole 32-byte sector in one request.  The direct epilogues used to write it
// as two 16-byte halves -- twice the L1 -> L2 write transactions (3.1 M per FFN-up / QKV launch at 1024 streams, profiles/r02_gemm...).
__device__ __forceinline__ void st_global_256(void* dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g,
                                              uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h)
               : "memory");
}
__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
}
// Ragged right edge (N = 8198) or output rows that are not 16-byte aligned: generic pairwise path, kept out of line so
// that the hot loop of every specialisation stays small enough for the instruction cache.
__device__ __noinline__ void epilogue_edge(const EpiParams& p, int m, int n, int N, float4 v) {
  epilogue_pair(p, m, n, N, v.x, v.y);
  if (n + 2 < N) epilogue_pair(p, m, n + 2, N, v.z, v.w);
}

// Fused epilogue on 4 adjacent columns n..n+3 (n % 4 == 0, n + 3 < N) of row m; MODE is a compile-time EpiMode.
// Per-row context, computed once per (tile, row) instead of once per quad: the mapped output row (row-map mode, < 0 =
// dropped) or the ring row slot * kRingCap + phys of the stream this packed row belongs to (QKV mode).
template <int MODE>
__device__ __forceinline__ int epilogue_row_ctx(const EpiParams& p, int m, int split) {
  if constexpr (MODE == EPI_PARTIAL_F32) {
    return split * p.part_rows + m;
  } else if constexpr (MODE == EPI_BIAS_ROWMAP_F32) {
    return p.row_map[m];
  } else if constexpr (MODE == EPI_QKV) {
    const int e = p.row_entry[m];
    int phys = p.entry_head[e] + kCacheS + p.row_pos[m];      // head in [0, kRingCap), row_pos in [-kCacheS, kMaxTq)
    phys -= phys >= kRingCap ? kRingCap : 0;
    // bf16 mode (k_natural): K and V rings are head-major [slot][head][kRingCap][128]; precise mode: K^T / V rings per slot
    return p.entry_slot[e] * (p.k_natural ? kHeads * kRingCap : kRingCap) + phys;
  } else {
    return m;
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_quad(const EpiParams& p, int m, int ctx, int n, float4 v) {
  const int nn = n + p.n_off;
  if constexpr (MODE == EPI_F32) {
    *reinterpret_cast<float4*>(p.out_f32 + (size_t)m * p.ldo + nn) = v;
  } else if constexpr (MODE == EPI_PARTIAL_F32) {
    if (p.part_bf16) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out_f32) + (size_t)ctx * p.ldo + nn) = pack4_bf16(v.x, v.y, v.z, v.w);
    else *reinterpret_cast<float4*>(p.out_f32 + (size_t)ctx * p.ldo + nn) = v;
  } else if constexpr (MODE == EPI_BIAS_F32 || MODE == EPI_BIAS_RELU_F32 || MODE == EPI_BIAS_ROWMAP_F32) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + nn);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if constexpr (MODE == EPI_BIAS_RELU_F32) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (ctx < 0) return;                                      // ctx == m except in row-map mode
    *reinterpret_cast<float4*>(p.out_f32 + (size_t)ctx * p.ldo + nn) = v;
  } else if constexpr (MODE == EPI_BIAS_RELU_ACT) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias + nn);
    v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f); v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
    store_act4(p.out_act, m, p.lda_out, nn, v, p.lo_off_out);
  } else if constexpr (MODE == EPI_ACT) {
    store_act4(p.out_act, m, p.lda_out, nn, v, p.lo_off_out);
  } else if constexpr (MODE == EPI_SILU_ACT) {
    v.x = silu(v.x); v.y = silu(v.y); v.z = silu(v.z); v.w = silu(v.w);
    store_act4(p.out_act, m, p.lda_out, nn, v, p.lo_off_out);
  } else if constexpr (MODE == EPI_RESADD_F32) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + (size_t)m * p.ldo + nn);
    float4 x = *o;
    x.x += p.scale * v.x; x.y += p.scale * v.y; x.z += p.scale * v.z; x.w += p.scale * v.w;
    *o = x;
  } else if constexpr (MODE == EPI_GLU_F32) {
    const float g0 = v.x * sigmoidf_(v.y), g1 = v.z * sigmoidf_(v.w);
    if (p.out_act) {      // bf16 mode: the depthwise kernel reads bf16
      const __nv_bfloat162 h = __floats2bfloat162_rn(g0, g1);
      *reinterpret_cast<uint32_t*>(p.out_act + (size_t)m * p.lda_out + (nn >> 1)) = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      *reinterpret_cast<float2*>(p.out_f32 + (size_t)m * p.ldo + (nn >> 1)) = make_float2(g0, g1);
    }
  } else if constexpr (MODE == EPI_QKV) {
    if (nn < kDModel) {
      if (p.q_bf16) {
        const float4 u = *reinterpret_cast<const float4*>(p.bias_u + nn), w = *reinterpret_cast<const float4*>(p.bias_v + nn);
        __nv_bfloat16* q = p.q_bf16 + (size_t)m * kDModel + nn;
        *reinterpret_cast<uint2*>(q) = pack4_bf16(v.x + u.x, v.y + u.y, v.z + u.z, v.w + u.w);
        *reinterpret_cast<uint2*>(q + p.q_plane) = pack4_bf16(v.x + w.x, v.y + w.y, v.z + w.z, v.w + w.w);
      } else {
        *reinterpret_cast<float4*>(p.out_f32 + (size_t)m * p.ldo + nn) = v;
      }
      return;
    }
    if (p.k_natural) {
      const int c = (nn - kDModel) & (kDModel - 1), h = c >> 7, d = c & 127;
      const size_t i0 = ((size_t)ctx + (size_t)h * kRingCap) * kDHead + d;
      void* ring = nn < 2 * kDModel ? p.kring : p.vring;
      if (p.kv_f32) *reinterpret_cast<float4*>((float*)ring + i0) = v;
      else *reinterpret_cast<uint2*>((__nv_bfloat16*)ring + i0) = pack4_bf16(v.x, v.y, v.z, v.w);
    } else if (nn < 2 * kDModel) {
      const int c = nn - kDModel, h = c >> 7, d = c & 127;
      const int slot = ctx / kRingCap, phys = ctx - slot * kRingCap;
      const size_t i0 = (((size_t)slot * kHeads + h) * kDHead + d) * kRingCap + phys;   // K^T ring: 4 rows kRingCap apart
      if (p.kv_f32) {
        float* k = (float*)p.kring + i0;
        k[0] = v.x; k[kRingCap] = v.y; k[2 * kRingCap] = v.z; k[3 * kRingCap] = v.w;
      } else {
        __nv_bfloat16* k = (__nv_bfloat16*)p.kring + i0;
        k[0] = __float2bfloat16_rn(v.x); k[kRingCap] = __float2bfloat16_rn(v.y);
        k[2 * kRingCap] = __float2bfloat16_rn(v.z); k[3 * kRingCap] = __float2bfloat16_rn(v.w);
      }
    } else {
      const size_t i0 = (size_t)ctx * kDModel + (nn - 2 * kDModel);
      if (p.kv_f32) *reinterpret_cast<float4*>((float*)p.vring + i0) = v;
      else *reinterpret_cast<uint2*>((__nv_bfloat16*)p.vring + i0) = pack4_bf16(v.x, v.y, v.z, v.w);
    }
  }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmArgs g,
               const int lo_row_off) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                                    // [kStages][16 KB]
  uint8_t* sB = base + C::kStages * kTileABytes;         // [kStages][BN*128 B]
  float* sEpi = reinterpret_cast<float*>(base + C::kStages * C::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sEpi) + kEpiWarps * kEpiBytesPerWarp);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // every full quad takes the vector path unless the output rows are not 16-byte aligned (f32 rows with ldo % 4 != 0)
  const bool f32_rows = MODE == EPI_F32 || MODE == EPI_BIAS_F32 || MODE == EPI_BIAS_RELU_F32 || MODE == EPI_BIAS_ROWMAP_F32 ||
                        MODE == EPI_RESADD_F32;
  const bool ragged = f32_rows && (g.epi.ldo & 3);
  const int kb_per_pass = g.K / BK;
  const int n_pass = g.a_lo_off != 0 ? 2 : 1;
  const int splits = MODE == EPI_PARTIAL_F32 ? g.epi.splits : 1;      // work unit = (tile, k-split), split fastest
  // bf16 rows (plain operand plane / bf16 partial sums) with 16-byte aligned 16-column groups: registers -> global directly
  constexpr bool kDirectMode = MODE == EPI_ACT || MODE == EPI_SILU_ACT || MODE == EPI_BIAS_RELU_ACT || MODE == EPI_PARTIAL_F32 ||
                               MODE == EPI_GLU_F32 || MODE == EPI_QKV;
  const bool direct = kDirectMode && g.epi.direct_bf16 != 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < C::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // barrier init + TMEM allocation above overlap the previous kernel's tail.  The weight tiles are constants: the producer requests
  // the W half of its first ring stages BEFORE waiting for the predecessor grid (the HBM round trip of this launch's weights overlaps
  // the previous launch; only the A tiles -- the predecessor's output -- wait).  Not when the row count is read from device memory
  // (decode GEMMs): the tile walk depends on it.
  pdl_trigger();
  int pre = 0;
  if (warp == 0 && lane == 0 && g.M_dev == nullptr) {
    const int tiles_m0 = (g.M + BM - 1) / BM, tiles_mn0 = tiles_m0 * ((g.N + BN - 1) / BN);
    const int unit = blockIdx.x;
    if (unit < tiles_mn0 * splits * g.batch) {
      const int tile_b = unit / splits, sp = unit - tile_b * splits;
      const int bt = tile_b / tiles_mn0, tile = tile_b - bt * tiles_mn0;
      const int n0 = (tile / tiles_m0) * BN;
      const int kb_lo = sp * kb_per_pass / splits, kbs = (sp + 1) * kb_per_pass / splits - kb_lo;
      const int num_kb = kbs * n_pass;
      const int w_row0 = bt * g.w_row_stride;
      pre = num_kb < C::kStages ? num_kb : C::kStages;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_expect_tx(&full_bar[kb], C::kStageBytes);
        const int kk = (kb_lo + kb % kbs) * BK;
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)
          tma_load_2d(sB + kb * C::kTileBBytes + j * kTileABytes, &map_w, &full_bar[kb], kk, w_row0 + n0 + j * 128);
      }
    }
  }
  pdl_wait();
  const int M = g.M_dev ? min(*g.M_dev, g.M) : g.M;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
  const int tiles_mn = tiles_m * tiles_n;
  const int total_tiles = tiles_mn * splits * g.batch;  // CTAs beyond the (device-side) unit count fall through to the teardown

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x) {
        const int tile_b = unit / splits, sp = unit - tile_b * splits;
        const int bt = tile_b / tiles_mn, tile = tile_b - bt * tiles_mn;
        const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
        const int kb_lo = sp * kb_per_pass / splits, kbs = (sp + 1) * kb_per_pass / splits - kb_lo;
        const int num_kb = kbs * n_pass;
        const int a_col0 = bt * g.a_col_stride, w_row0 = bt * g.w_row_stride;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          const bool w_done = it < pre;      // stage armed and its W tiles requested before the dependency wait
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (!w_done) mbar_expect_tx(&full_bar[s], C::kStageBytes);
          const int pass = kb / kbs, kk = (kb_lo + kb % kbs) * BK;
          tma_load_2d(sA + s * kTileABytes, &map_a, &full_bar[s], a_col0 + kk, m0 + pass * lo_row_off);
          if (!w_done) {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              tma_load_2d(sB + s * C::kTileBBytes + j * kTileABytes, &map_w, &full_bar[s], kk, w_row0 + n0 + j * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0, tl = 0;
      for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x, ++tl) {
        const int sp = unit % splits;
        const int num_kb = ((sp + 1) * kb_per_pass / splits - sp * kb_per_pass / splits) * n_pass;
        const int acc = tl & 1;
        mbar_wait(&tmem_empty_bar[acc], ((tl >> 1) & 1) ^ 1);          // epilogue has drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(smem_u32(sA + s * kTileABytes));
          const uint64_t db = make_smem_desc(smem_u32(sB + s * C::kTileBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)                  // +32 bytes per K=16 step inside the 128-byte swizzle atom
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDesc<BN>::value, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);                       // smem slot reusable once these MMAs have read it
        }
        umma_commit(&tmem_full_bar[acc]);                   // accumulator complete
      }
    }
  } else {
    // epilogue warps 2..9: TMEM lane quarter = warp % 4 (hardware rule), column half = (warp - 2) / 4
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stage = sEpi + (warp - 2) * (32 * kEpiPitch);
    const int rsel = lane >> 2, c4 = lane & 3;
    const int rd_row = (rsel & 1) * 4 + (rsel >> 1);        // rows r and r+4 in one quarter-warp: conflict-free reads
    int tl = 0;
    for (int unit = blockIdx.x; unit < total_tiles; unit += gridDim.x, ++tl) {
      const int tile_b = unit / splits, sp = unit - tile_b * splits;
      const int bt = tile_b / tiles_mn, tile = tile_b - bt * tiles_mn;
      const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;
      const int n_out0 = bt * g.out_col_stride;      // batched problems: column offset of this batch in the output
      const int acc = tl & 1;
      mbar_wait(&tmem_full_bar[acc], (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
      int ctx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + q * 32 + rd_row + 8 * i;
        ctx[i] = m < M ? epilogue_row_ctx<MODE>(g.epi, m, sp) : -1;
      }
      float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // EPI_ARGMAX: running first-maximum per row
      int bidx[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
      // direct path (bf16 outputs): the lane keeps its own TMEM row, so no staging tile is involved
      const int m_d = m0 + q * 32 + lane;
      __nv_bfloat16* drow = nullptr;
      if (direct && m_d < M) {
        if constexpr (MODE == EPI_PARTIAL_F32)
          drow = reinterpret_cast<__nv_bfloat16*>(g.epi.out_f32) + (size_t)epilogue_row_ctx<MODE>(g.epi, m_d, sp) * g.epi.ldo;
        else if constexpr (MODE == EPI_QKV)
          drow = g.epi.q_bf16 + (size_t)m_d * kDModel;
        else
          drow = g.epi.out_act + (size_t)m_d * g.epi.lda_out;
      }
      if constexpr (MODE == EPI_QKV) {
        if (direct) {
          // bf16 streaming mode: q leaves as the two biased planes, k and v go to the head-major rings, all straight from the TMEM-load
          // registers -- a lane keeps its accumulator row, so each 32-column chunk is 64 contiguous bytes of one destination row
          // (the staged path scattered 8-byte pieces through the smem transpose: 46 - 50 us for this launch at 1024 streams)
          const int ctxd = m_d < M ? epilogue_row_ctx<MODE>(g.epi, m_d, sp) : 0;
#pragma unroll 1
          for (int c0 = 0; c0 < BN / 2; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            if (c0 + 32 == BN / 2) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            const int ncol = n0 + half * (BN / 2) + c0;
            if (drow == nullptr || ncol >= g.N) continue;
            if (ncol < kDModel) {
              uint4* du = reinterpret_cast<uint4*>(drow + ncol);
              uint4* dv = reinterpret_cast<uint4*>(drow + g.epi.q_plane + ncol);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                uint2 pa[4], pb[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int c = 16 * jj + 4 * x;
                  const float4 u = *reinterpret_cast<const float4*>(g.epi.bias_u + ncol + c), w = *reinterpret_cast<const float4*>(g.epi.bias_v + ncol + c);
                  const float f0 = __uint_as_float(v[c]), f1 = __uint_as_float(v[c + 1]), f2 = __uint_as_float(v[c + 2]), f3 = __uint_as_float(v[c + 3]);
                  pa[x] = pack4_bf16(f0 + u.x, f1 + u.y, f2 + u.z, f3 + u.w);
                  pb[x] = pack4_bf16(f0 + w.x, f1 + w.y, f2 + w.z, f3 + w.w);
                }
                st_global_256(du + 2 * jj, pa[0].x, pa[0].y, pa[1].x, pa[1].y, pa[2].x, pa[2].y, pa[3].x, pa[3].y);
                st_global_256(dv + 2 * jj, pb[0].x, pb[0].y, pb[1].x, pb[1].y, pb[2].x, pb[2].y, pb[3].x, pb[3].y);
              }
            } else {
              const int c = (ncol - kDModel) & (kDModel - 1), h = c >> 7, d = c & 127;
              __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(ncol < 2 * kDModel ? g.epi.kring : g.epi.vring);
              uint4* dst = reinterpret_cast<uint4*>(ring + ((size_t)ctxd + (size_t)h * kRingCap) * kDHead + d);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                uint2 pa[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int c = 16 * jj + 4 * x;
                  pa[x] = pack4_bf16(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
                }
                st_global_256(dst + 2 * jj, pa[0].x, pa[0].y, pa[1].x, pa[1].y, pa[2].x, pa[2].y, pa[3].x, pa[3].y);
              }
            }
          }
          continue;
        }
      }
      if constexpr (MODE == EPI_GLU_F32) {
        if (direct) {
          // GLU straight from the TMEM-load registers: 32 accumulator columns = 16 (value, gate) pairs of one row per lane ->
          // 16 bf16 outputs = one full 32-byte sector, no shared-memory transpose (the transposed path made this launch
          // epilogue-bound: 46 us for 25.8 GFLOP at 1024 streams)
#pragma unroll 1
          for (int c0 = 0; c0 < BN / 2; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            if (c0 + 32 == BN / 2) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            const int ncol = n0 + half * (BN / 2) + c0;
            if (drow != nullptr && ncol < g.N) {
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a0 = __uint_as_float(v[4 * j]) * sigmoidf_(__uint_as_float(v[4 * j + 1]));
                const float a1 = __uint_as_float(v[4 * j + 2]) * sigmoidf_(__uint_as_float(v[4 * j + 3]));
                const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
                o[j] = *reinterpret_cast<const uint32_t*>(&h);
              }
              uint4* dst = reinterpret_cast<uint4*>(drow + (ncol >> 1));
              st_global_256(dst, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
            }
          }
          continue;
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN / 2; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        if (c0 + 16 == BN / 2) {                            // last TMEM read of this tile: give the buffer back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (direct) {
          // 16 adjacent columns of one row per lane -> 32 contiguous bytes of bf16 = one full sector, written straight from the
          // registers the TMEM load filled (no shared-memory transpose: its ~50 extra instructions per chunk made short-K
          // GEMMs epilogue-bound)
          const int nn = n0 + half * (BN / 2) + c0 + n_out0 + g.epi.n_off;
          if (drow != nullptr && n0 + half * (BN / 2) + c0 < g.N) {
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if constexpr (MODE == EPI_SILU_ACT) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = silu(f[j]);
            } else if constexpr (MODE == EPI_BIAS_RELU_ACT) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 bq = *reinterpret_cast<const float4*>(g.epi.bias + nn + 4 * j);
                f[4 * j] = fmaxf(f[4 * j] + bq.x, 0.f); f[4 * j + 1] = fmaxf(f[4 * j + 1] + bq.y, 0.f);
                f[4 * j + 2] = fmaxf(f[4 * j + 2] + bq.z, 0.f); f[4 * j + 3] = fmaxf(f[4 * j + 3] + bq.w, 0.f);
              }
            }
            const uint2 p0 = pack4_bf16(f[0], f[1], f[2], f[3]), p1 = pack4_bf16(f[4], f[5], f[6], f[7]);
            const uint2 p2 = pack4_bf16(f[8], f[9], f[10], f[11]), p3 = pack4_bf16(f[12], f[13], f[14], f[15]);
            uint4* dst = reinterpret_cast<uint4*>(drow + nn);
            st_global_256(dst, p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y);
          }
          continue;
        }
        __syncwarp();                                       // previous chunk's readers are done with the staging tile
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(stage + lane * kEpiPitch + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const int n = n0 + half * (BN / 2) + c0 + 4 * c4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rd_row + 8 * i;
          const float4 val = *reinterpret_cast<const float4*>(stage + r * kEpiPitch + 4 * c4);
          const int m = m0 + q * 32 + r;
          if constexpr (MODE == EPI_ARGMAX) {
            if (m < M && n < g.N) {
              const float vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                const int col = n + x;
                if (col >= g.N) break;
                float y = vv[x] + g.epi.bias[col];
                if (y != y) y = -100.0f;                                  // NaN logits -> -100 (parakeet_trt.cpp:2971)
                if (col == kBlank) y -= g.epi.blank_penalty;              // PARAKEET_BLANK_PENALTY (:3175-3178)
                if (col < kVocab) { if (y > best[i]) { best[i] = y; bidx[i] = col; } }   // ascending columns: first maximum wins
                else g.epi.dur_out[(size_t)m * kNDur + (col - kVocab)] = y;
              }
            }
          } else {
            if (m < M && n < g.N) {
              if (n + 3 < g.N && !ragged) epilogue_quad<MODE>(g.epi, m, ctx[i], n + n_out0, val);
              else epilogue_edge(g.epi, m, n + n_out0, g.N + n_out0, val);
            }
          }
        }
      }
      if constexpr (MODE == EPI_ARGMAX) {
        // the 4 lanes c4 = 0..3 of a row hold disjoint column sets: warp-level (value desc, index asc) reduction, then one
        // (max, argmax) per row and 128-column slab
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int off = 1; off <= 2; off <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best[i], off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[i], off);
            if (ov > best[i] || (ov == best[i] && oi < bidx[i])) { best[i] = ov; bidx[i] = oi; }
          }
          const int m = m0 + q * 32 + rd_row + 8 * i;
          if (c4 == 0 && m < M) {
            const size_t o = (size_t)m * kArgmaxParts + (size_t)(n0 / BN) * 2 + half;
            g.epi.part_val[o] = best[i];
            g.epi.part_idx[o] = bidx[i];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
  }
}

// ================================================================================================ 2-CTA variant
// CTA pair (cluster of 2, tcgen05 cta_group::2) computing one 256 x 256 tile: CTA r of the pair stages the A rows
// [m0 + 128 r, +128) and the W rows [n0 + 128 r, +128) of every k-block (32 KB per CTA per k-block for 128 x 256 x 64 MACs
// per SM -- half the L2->SM operand traffic per MAC of the 128 x 128 single-CTA tile, which is what bounds that kernel).
// The leader CTA (rank 0) issues tcgen05.mma.cta_group::2 (M = 256, N = 256); each SM accumulates its own 128 rows x 256
// columns in its own TMEM (double-buffered: 512 columns) and runs its own epilogue.
//   full[s]       : leader only; both CTAs' TMA loads complete_tx on it (64 KB per stage)
//   empty[s]      : one per CTA; released by the leader's multicast tcgen05.commit
//   tmem_full[a]  : one per CTA; multicast commit after the last k-block of a tile
//   tmem_empty[a] : leader only; 16 arrivals (8 epilogue warps x 2 CTAs, the peer's arrive remotely)
constexpr int kStages2 = 6;
constexpr int kStageBytes2 = 2 * kTileABytes;
constexpr size_t kSmemBytes2 = 1024 + (size_t)kStages2 * kStageBytes2 + (size_t)kEpiWarps * kEpiBytesPerWarp + 256;
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in CTA rank 0 of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc_v, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc_v), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {      // arrives on `bar` at the same offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_w32, const GemmArgs g, const int lo_row_off, const int tail_c, const int full_units) {
  constexpr int BN = 256, BM2 = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                                    // [kStages2][16 KB]  this CTA's 128 A rows
  uint8_t* sB = base + kStages2 * kTileABytes;           // [kStages2][16 KB]  this CTA's 128 W rows
  float* sEpi = reinterpret_cast<float*>(base + kStages2 * kStageBytes2);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sEpi) + kEpiWarps * kEpiBytesPerWarp);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full_bar = empty_bar + kStages2;        // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const bool f32_rows = MODE == EPI_F32 || MODE == EPI_BIAS_F32 || MODE == EPI_BIAS_RELU_F32 || MODE == EPI_BIAS_ROWMAP_F32 ||
                        MODE == EPI_RESADD_F32;
  const bool ragged = f32_rows && (g.epi.ldo & 3);
  const int kb_per_pass = g.K / BK;
  const int n_pass = g.a_lo_off != 0 ? 2 : 1;
  const int splits = MODE == EPI_PARTIAL_F32 ? g.epi.splits : 1;      // work unit = (pair tile, k-split), split fastest

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                    // barrier inits and the TMEM allocation are visible in both CTAs
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // (as in the single-CTA kernel: the W halves of the first ring stages are requested before the dependency wait)
  pdl_trigger();
  int pre = 0;
  if (warp == 0 && lane == 0 && g.M_dev == nullptr) {
    const int tiles_m0 = (g.M + BM2 - 1) / BM2;
    if (pair < tiles_m0 * ((g.N + BN - 1) / BN) * splits) {
      const int tile = pair / splits, sp = pair - tile * splits;
      const int n0 = (tile / tiles_m0) * BN + 128 * rank;
      const int kb_lo = sp * kb_per_pass / splits, kbs = (sp + 1) * kb_per_pass / splits - kb_lo;
      const int num_kb = kbs * n_pass;
      pre = num_kb < kStages2 ? num_kb : kStages2;
      for (int kb = 0; kb < pre; ++kb) {
        if (rank == 0) mbar_expect_tx(&full_bar[kb], 2 * kStageBytes2);
        tma_load_2d_2sm(sB + kb * kTileABytes, &map_w, smem_u32(&full_bar[kb]) & kPeerBitMask, (kb_lo + kb % kbs) * BK, n0);
      }
    }
  }
  pdl_wait();
  const int M = g.M_dev ? min(*g.M_dev, g.M) : g.M;
  const int tiles_m = (M + BM2 - 1) / BM2, tiles_n = (g.N + BN - 1) / BN;
  // Work units.  v < full_units: one 256 x 256 tile (or k-split of one).  The tiles of the last, partly filled round are cut into
  // tail_c column slices of 256 / tail_c columns each (tail_c = 1, 2 or 4, chosen by the host so that the slices fill the pairs
  // better than the whole tiles would): v >= full_units walks (tile, slice) with the slice index fastest.  full_units is a multiple
  // of the pair count, so `v += n_pairs` runs from the full rounds straight into the tail.
  const int total_tiles = tail_c > 1 ? full_units + (tiles_m * tiles_n * splits - full_units) * tail_c : tiles_m * tiles_n * splits;
  const int tail_shift = tail_c == 4 ? 2 : tail_c == 2 ? 1 : 0;
  auto decode = [&](int v, int& unit, int& bn, int& ncol) {
    if (v < full_units || tail_c <= 1) { unit = v; bn = BN; ncol = 0; }
    else { const int sub = v - full_units; unit = full_units + (sub >> tail_shift); bn = BN >> tail_shift; ncol = (sub & (tail_c - 1)) * bn; }
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int v = pair; v < total_tiles; v += n_pairs) {
        int unit, bn, ncol;
        decode(v, unit, bn, ncol);
        const int tile = unit / splits, sp = unit - tile * splits;
        const int m0 = (tile % tiles_m) * BM2 + 128 * rank;
        const int n0 = (tile / tiles_m) * BN + ncol + (bn >> 1) * rank;      // this CTA stages bn / 2 of the unit's W rows
        const int kb_lo = sp * kb_per_pass / splits, kbs = (sp + 1) * kb_per_pass / splits - kb_lo;
        const int num_kb = kbs * n_pass;
        const uint32_t stage_tx = 2u * (uint32_t)(kTileABytes + (bn >> 1) * 128);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages2;
          const uint32_t ph = (it / kStages2) & 1;
          const bool w_done = it < pre;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0 && !w_done) mbar_expect_tx(&full_bar[s], stage_tx);
          const uint32_t leader_full = smem_u32(&full_bar[s]) & kPeerBitMask;
          const int pass = kb / kbs, kk = (kb_lo + kb % kbs) * BK;
          tma_load_2d_2sm(sA + s * kTileABytes, &map_a, leader_full, kk, m0 + pass * lo_row_off);
          if (bn == BN) {
            if (!w_done) tma_load_2d_2sm(sB + s * kTileABytes, &map_w, leader_full, kk, n0);
          } else {      // column slice: 64 or 32 W rows per CTA through the 32-row box map
            for (int j = 0; j < (bn >> 6); ++j) tma_load_2d_2sm(sB + s * kTileABytes + j * 4096, &map_w32, leader_full, kk, n0 + 32 * j);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int it = 0, tl = 0;
      for (int v = pair; v < total_tiles; v += n_pairs, ++tl) {
        int unit, bn, ncol;
        decode(v, unit, bn, ncol);
        const int sp = unit % splits;
        const int num_kb = ((sp + 1) * kb_per_pass / splits - sp * kb_per_pass / splits) * n_pass;
        const uint32_t idesc = (kIdesc2 & ~(0x3Fu << 17)) | ((uint32_t)(bn >> 3) << 17);      // N = bn
        const int acc = tl & 1;
        mbar_wait(&tmem_empty_bar[acc], ((tl >> 1) & 1) ^ 1);          // both CTAs' epilogues have drained this buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages2;
          const uint32_t ph = (it / kStages2) & 1;
          mbar_wait(&full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc(smem_u32(sA + s * kTileABytes));
          const uint64_t db = make_smem_desc(smem_u32(sB + s * kTileABytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[s]);
        }
        umma_commit_2sm(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stage = sEpi + (warp - 2) * (32 * kEpiPitch);
    const int rsel = lane >> 2, c4 = lane & 3;
    const int rd_row = (rsel & 1) * 4 + (rsel >> 1);
    int tl = 0;
    for (int v = pair; v < total_tiles; v += n_pairs, ++tl) {
      int unit, bn, ncol_u;
      decode(v, unit, bn, ncol_u);
      const int hb = bn >> 1;                                // columns per epilogue half (128, 64 or 32)
      const int tile = unit / splits, sp = unit - tile * splits;
      const int m0 = (tile % tiles_m) * BM2 + 128 * rank, n0 = (tile / tiles_m) * BN + ncol_u;
      const int acc = tl & 1;
      mbar_wait(&tmem_full_bar[acc], (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * hb);
      int ctx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + q * 32 + rd_row + 8 * i;
        ctx[i] = m < M ? epilogue_row_ctx<MODE>(g.epi, m, sp) : -1;
      }
      if constexpr (MODE == EPI_GLU_F32 || MODE == EPI_QKV) {
        if (g.epi.direct_bf16) {      // same direct epilogues as the single-CTA kernel (see there)
          const int m_d = m0 + q * 32 + lane;
          const bool row_ok = m_d < M;
          const int ctxd = (MODE == EPI_QKV && row_ok) ? epilogue_row_ctx<MODE>(g.epi, m_d, sp) : 0;
#pragma unroll 1
          for (int c0 = 0; c0 < hb; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            if (c0 + 32 == hb) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive_rank0(&tmem_empty_bar[acc]);
            }
            const int ncol = n0 + half * hb + c0;
            if (!row_ok || ncol >= g.N) continue;
            if constexpr (MODE == EPI_GLU_F32) {
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a0 = __uint_as_float(v[4 * j]) * sigmoidf_(__uint_as_float(v[4 * j + 1]));
                const float a1 = __uint_as_float(v[4 * j + 2]) * sigmoidf_(__uint_as_float(v[4 * j + 3]));
                const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
                o[j] = *reinterpret_cast<const uint32_t*>(&hh);
              }
              uint4* dst = reinterpret_cast<uint4*>(g.epi.out_act + (size_t)m_d * g.epi.lda_out + (ncol >> 1));
              st_global_256(dst, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
            } else if (ncol < kDModel) {
              __nv_bfloat16* drow = g.epi.q_bf16 + (size_t)m_d * kDModel;
              uint4* du = reinterpret_cast<uint4*>(drow + ncol);
              uint4* dv = reinterpret_cast<uint4*>(drow + g.epi.q_plane + ncol);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                uint2 pa[4], pb[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int c = 16 * jj + 4 * x;
                  const float4 u = *reinterpret_cast<const float4*>(g.epi.bias_u + ncol + c), w = *reinterpret_cast<const float4*>(g.epi.bias_v + ncol + c);
                  const float f0 = __uint_as_float(v[c]), f1 = __uint_as_float(v[c + 1]), f2 = __uint_as_float(v[c + 2]), f3 = __uint_as_float(v[c + 3]);
                  pa[x] = pack4_bf16(f0 + u.x, f1 + u.y, f2 + u.z, f3 + u.w);
                  pb[x] = pack4_bf16(f0 + w.x, f1 + w.y, f2 + w.z, f3 + w.w);
                }
                st_global_256(du + 2 * jj, pa[0].x, pa[0].y, pa[1].x, pa[1].y, pa[2].x, pa[2].y, pa[3].x, pa[3].y);
                st_global_256(dv + 2 * jj, pb[0].x, pb[0].y, pb[1].x, pb[1].y, pb[2].x, pb[2].y, pb[3].x, pb[3].y);
              }
            } else {
              const int c = (ncol - kDModel) & (kDModel - 1), h = c >> 7, d = c & 127;
              __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(ncol < 2 * kDModel ? g.epi.kring : g.epi.vring);
              uint4* dst = reinterpret_cast<uint4*>(ring + ((size_t)ctxd + (size_t)h * kRingCap) * kDHead + d);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                uint2 pa[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                  const int c = 16 * jj + 4 * x;
                  pa[x] = pack4_bf16(__uint_as_float(v[c]), __uint_as_float(v[c + 1]), __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
                }
                st_global_256(dst + 2 * jj, pa[0].x, pa[0].y, pa[1].x, pa[1].y, pa[2].x, pa[2].y, pa[3].x, pa[3].y);
              }
            }
          }
          continue;
        }
      }
      if constexpr (MODE == EPI_PARTIAL_F32 || MODE == EPI_SILU_ACT || MODE == EPI_ACT) {
        if (g.epi.direct_bf16) {
          // bf16 rows straight from the TMEM-load registers: the lane keeps its own accumulator row, 32 columns per load
          // = 64 contiguous bytes of the destination row (see the single-CTA kernel's direct path)
          const int m_d = m0 + q * 32 + lane;
          __nv_bfloat16* drow = nullptr;
          if (m_d < M) {
            if constexpr (MODE == EPI_PARTIAL_F32)
              drow = reinterpret_cast<__nv_bfloat16*>(g.epi.out_f32) + (size_t)epilogue_row_ctx<MODE>(g.epi, m_d, sp) * g.epi.ldo;
            else
              drow = g.epi.out_act + (size_t)m_d * g.epi.lda_out;
          }
#pragma unroll 1
          for (int c0 = 0; c0 < hb; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)c0, v);
            if (c0 + 32 == hb) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive_rank0(&tmem_empty_bar[acc]);
            }
            const int ncol = n0 + half * hb + c0;
            if (drow != nullptr && ncol < g.N) {
              uint4* dst = reinterpret_cast<uint4*>(drow + ncol + g.epi.n_off);
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                float f[16];
#pragma unroll
                for (int x = 0; x < 16; ++x) {
                  f[x] = __uint_as_float(v[16 * jj + x]);
                  if constexpr (MODE == EPI_SILU_ACT) f[x] = silu(f[x]);
                }
                const uint2 p0 = pack4_bf16(f[0], f[1], f[2], f[3]), p1 = pack4_bf16(f[4], f[5], f[6], f[7]);
                const uint2 p2 = pack4_bf16(f[8], f[9], f[10], f[11]), p3 = pack4_bf16(f[12], f[13], f[14], f[15]);
                st_global_256(dst + 2 * jj, p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y);
              }
            }
          }
          continue;
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < hb; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        if (c0 + 16 == hb) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_rank0(&tmem_empty_bar[acc]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(stage + lane * kEpiPitch + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        __syncwarp();
        const int n = n0 + half * hb + c0 + 4 * c4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rd_row + 8 * i;
          const float4 val = *reinterpret_cast<const float4*>(stage + r * kEpiPitch + 4 * c4);
          const int m = m0 + q * 32 + r;
          if (m < M && n < g.N) {
            if (n + 3 < g.N && !ragged) epilogue_quad<MODE>(g.epi, m, ctx[i], n, val);
            else epilogue_edge(g.epi, m, n, g.N, val);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                    // neither CTA may exit (or free TMEM) while its peer can still reach it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PKB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    PKB_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int g_two_cta = -1;       // PARAKEET_B200_GEMM_2CTA: 1 (default) = CTA-pair kernel where it applies, 0 = single-CTA kernel only
int g_force_bn = -1;      // 0: heuristic; 128 / 256: forced tile width (PARAKEET_B200_GEMM_BN or gemm_tc_set_bn)
int g_sm_count = 0;
int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    PKB_CUDA(cudaGetDevice(&dev));
    PKB_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  return g_sm_count;
}

// Tile width.  The kernel is bound by L2->SM operand delivery (~6.2 KB/clk chip-wide, profiles/): per k-block a 128x128 tile
// pulls 32 KB, a 128x256 tile 48 KB for twice the math, so a wide tile costs ~1.5 narrow ones; against that, narrow tiles
// quantise better into waves of `sms` CTAs.  Pick the smaller (waves x tile cost); narrow on ties (shorter tail).
int pick_bn(int M, int N, int sms) {
  const long long tm = (M + BM - 1) / BM;
  const long long t128 = tm * ((N + 127) / 128), t256 = tm * ((N + 255) / 256);
  const long long w128 = (t128 + sms - 1) / sms, w256 = (t256 + sms - 1) / sms;
  return (w256 * 3 < w128 * 2) ? 256 : 128;
}

// CTA-pair kernel or single-CTA kernel?  Measured on B200 (tools/kbench.cu, M = 6144): the pair kernel wins where its
// 256 x 256 tiles fill whole waves of sms/2 pairs (N = 3072: +11 %) and on long-K problems, which are the most bound by
// L2->SM operand delivery (N = 1024, K = 4096: +8 %); it loses where the coarser tiles quantise worse.
bool pick_two_cta(int M, int N, int K, int sms) {
  if (N % 256 != 0 || M < 2048) return false;
  const int pairs = sms / 2;
  const long long t2 = (long long)((M + 255) / 256) * (N / 256);
  const double eff2 = (double)t2 / (double)(((t2 + pairs - 1) / pairs) * pairs);
  const int bn = pick_bn(M, N, sms);
  const long long t1 = (long long)((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  const double eff1 = (double)t1 / (double)(((t1 + sms - 1) / sms) * sms);
  // (N = 1024, K = 1024, the attention-output and conv pointwise_conv2 projections: 16.9 vs 19.0 us although its 96 pair tiles
  // fill only 65 % of two rounds -- the 128 x 128 tiles of the single-CTA kernel are bound by operand delivery there)
  // (from 3072 rows = 512 streams: forcing the pair kernel for every projection takes the step 9.40 -> 8.82 ms at 512 streams and
  // 11.09 -> 10.44 ms at 640, and loses below 2560 rows and, for QKV / GLU, above 4096: gpurun r2y)
  return eff2 > eff1 + 0.05 || (K >= 4096 && M >= 3072) || (N == 1024 && K == 1024 && M >= 3072);
}

}  // namespace

void make_tensor_map_2d(TensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows) {
  static_assert(sizeof(CUtensorMap) == sizeof(TensorMap), "CUtensorMap size");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = get_encode_fn()(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PKB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

template <int BN, int MODE>
void launch_cfg(int grid, const CUtensorMap& ma, const CUtensorMap& mw, const GemmArgs& g, int lo_row_off, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    PKB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN>::kSmemBytes));
    attr = true;
  }
  launch_k(gemm_tc_kernel<BN, MODE>, dim3(grid), dim3(kThreads), Cfg<BN>::kSmemBytes, st, ma, mw, g, lo_row_off);
}

template <int MODE>
void launch_cfg2(int pairs, const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mw32, const GemmArgs& g, int lo_row_off, int tail_c,
                 int full_units, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    PKB_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2));
    attr = true;
  }
  // cluster size comes from __cluster_dims__; the launch adds the programmatic-dependent-launch attribute
  launch_k(gemm_tc2_kernel<MODE>, dim3(2 * pairs), dim3(kThreads), kSmemBytes2, st, ma, mw, mw32, g, lo_row_off, tail_c, full_units);
}

void gemm_tc_set_bn(int bn) { g_force_bn = bn; }

// Residual GEMMs on the CTA-pair kernel: 256 x 256 tiles can leave the last wave of sms/2 pairs mostly empty (M = 6144,
// N = 1024: 96 tiles on 74 pairs).  Returns the k-split count (1 or 2) that fills the waves better, 0 if the pair kernel would
// not be chosen for this shape at all.
int gemm_tc_pair_splits(int M, int N, int K) {
  const int sms = sm_count();
  if (g_two_cta < 0) { const char* v = getenv("PARAKEET_B200_GEMM_2CTA"); g_two_cta = v ? atoi(v) : 1; }
  if (g_force_bn > 0 || g_two_cta == 0 || !pick_two_cta(M, N, K, sms)) return 0;
  const int pairs = sms / 2;
  const long long t = (long long)((M + 255) / 256) * (N / 256);
  auto waves_cost = [&](int s) { return (double)((t * s + pairs - 1) / pairs) / s; };      // rounds x (1/s of the k-loop)
  return (K / BK >= 32 && waves_cost(2) < 0.85 * waves_cost(1)) ? 2 : 1;
}

bool gemm_tc_supported(const GemmArgs& g) {
  // (batched problems address a column window of A through the tensor map: lda may exceed K)
  return g.K % BK == 0 && (g.lda == g.K || (g.batch > 1 && g.a_lo_off == 0)) && (g.a_lo_off % g.lda) == 0 && g.M > 0 && g.N > 0;
}

void gemm_tc(const GemmArgs& g_in, const TensorMap& map_a, const TensorMap& map_w, cudaStream_t st) {
  PKB_CHECK(gemm_tc_supported(g_in), "gemm_tc: unsupported shape");
  GemmArgs g = g_in;
  {
    // direct bf16 epilogue (single-CTA kernel): every 16-column group of an output row must be one aligned 32-byte run
    static const bool allow = [] { const char* v = getenv("PARAKEET_B200_EPI_DIRECT"); return !(v && v[0] == '0'); }();
    const EpiParams& e = g.epi;
    // (every 16-column group of an output row is written as ONE 256-bit store: 32-byte aligned rows and column offsets)
    bool ok = allow && g.N % 16 == 0 && e.n_off % 16 == 0 && g.out_col_stride % 16 == 0;
    if (e.mode == EPI_PARTIAL_F32) ok = ok && e.part_bf16 && e.ldo % 16 == 0 && ((uintptr_t)e.out_f32 & 31) == 0 && g.N % 32 == 0;
    else if (e.mode == EPI_QKV)
      ok = ok && e.k_natural && !e.kv_f32 && e.q_bf16 != nullptr && e.n_off == 0 && g.batch == 1 && g.N == 3 * kDModel &&
           ((uintptr_t)e.q_bf16 & 31) == 0 && ((uintptr_t)e.kring & 31) == 0 && ((uintptr_t)e.vring & 31) == 0 && (e.q_plane & 15) == 0 &&
           ((uintptr_t)e.bias_u & 15) == 0 && ((uintptr_t)e.bias_v & 15) == 0;
    else if (e.mode == EPI_GLU_F32)
      ok = ok && e.out_act != nullptr && g.N % 32 == 0 && e.n_off == 0 && g.batch == 1 && e.lda_out % 16 == 0 && ((uintptr_t)e.out_act & 31) == 0;
    else if (e.mode == EPI_ACT || e.mode == EPI_SILU_ACT || e.mode == EPI_BIAS_RELU_ACT)
      ok = ok && e.lo_off_out == 0 && e.lda_out % 16 == 0 && ((uintptr_t)e.out_act & 31) == 0 && (e.mode != EPI_BIAS_RELU_ACT || ((uintptr_t)e.bias & 15) == 0);
    else ok = false;
    g.epi.direct_bf16 = ok ? 1 : 0;
  }
  const int sms = sm_count();
  if (g_force_bn < 0) { const char* v = getenv("PARAKEET_B200_GEMM_BN"); g_force_bn = v ? atoi(v) : 0; }
  if (g_two_cta < 0) { const char* v = getenv("PARAKEET_B200_GEMM_2CTA"); g_two_cta = v ? atoi(v) : 1; }
  // PARAKEET_B200_PAIR_MODES: bit m set = epilogue mode m always takes the CTA-pair kernel when its shape allows.  Default: the FFN-up
  // projections (EPI_SILU_ACT, N = 4096, K = 1024) -- with the direct bf16 epilogue the 256 x 256 pair tile (32 KB of operands per SM
  // and k-block instead of 48 KB for the same MACs) wins although both tilings leave the last of 6 waves equally empty: 55.3 -> ~50 us
  // per launch, 17.48 -> 17.27 ms per step at 1024 streams (A/B in gpurun r2k; 0 restores the single-CTA choice)
  static const int pair_modes = [] { const char* v = getenv("PARAKEET_B200_PAIR_MODES"); return v ? atoi(v) : (1 << EPI_SILU_ACT); }();
  static const int pair_min_m = [] { const char* v = getenv("PARAKEET_B200_PAIR_MIN_M"); return v ? atoi(v) : 2048; }();
  const bool pair_forced = ((pair_modes >> g.epi.mode) & 1) && g.N % 256 == 0 && g.M >= pair_min_m && g.N % 32 == 0;
  const bool two_cta = g.epi.mode != EPI_ARGMAX && g.epi.mode != EPI_ACT && g.batch == 1 && !(g.epi.mode == EPI_PARTIAL_F32 && g.epi.splits > 1 && !g.epi.pair_split) &&
                       (g_force_bn == 512 || (g_force_bn == 0 && g_two_cta != 0 &&
                                              (pair_forced || (g.epi.mode == EPI_PARTIAL_F32 && g.epi.pair_split == 2) || pick_two_cta(g.M, g.N, g.K, sms))));
  if (two_cta) {
    const int pair_tiles = ((g.M + 255) / 256) * ((g.N + 255) / 256) * (g.epi.mode == EPI_PARTIAL_F32 ? g.epi.splits : 1);
    // (PARAKEET_B200_GEMM_MAX_PAIRS: measurement aid -- fewer resident pairs show how much of a tile's time is L2 -> SM contention)
    static const int max_pairs = [] { const char* v = getenv("PARAKEET_B200_GEMM_MAX_PAIRS"); return v ? atoi(v) : 1 << 30; }();
    const int pair_cap = sms / 2 < max_pairs ? sms / 2 : max_pairs;
    const int pairs = pair_tiles < pair_cap ? pair_tiles : pair_cap;
    const int lo_row_off2 = (int)(g.a_lo_off / g.lda);
    const CUtensorMap& ma2 = *reinterpret_cast<const CUtensorMap*>(&map_a);
    const CUtensorMap& mw2 = *reinterpret_cast<const CUtensorMap*>(&map_w);
    // Tail: the tiles of the last, partly filled round are cut into 2 or 4 column slices when that shortens the round
    // (cost of a tail cut c ways = ceil(r c / pairs) / c full-tile times).  Needs the weight's 32-row box map.
    // A slice is not proportionally cheaper than the tile: every slice re-reads the tile's 256 A rows and a narrow UMMA (N = 128 /
    // 64) is bound by operand delivery, so a round of half-width slices costs ~0.6 and a round of quarter-width slices ~0.45 of a
    // full-tile round (PARAKEET_B200_GEMM_TAIL_T2 / _T4, per mille; measured with tools/kbench.cu).  PARAKEET_B200_GEMM_TAIL = largest
    // cut allowed (0 / 1: never, 2, 4).
    // Measured (gpurun r2u, M = 6144): only the FFN-up shape gains (39.2 -> 38.0 us); k-split, GLU and QKV shapes lose 8 - 25 % and the
    // 1024-stream step does not move (15.64 vs 15.65 ms) -- so the cut is OFF by default and kept as a tested option.
    static const int tail_max = [] { const char* v = getenv("PARAKEET_B200_GEMM_TAIL"); return v ? atoi(v) : 0; }();
    static const double t2 = [] { const char* v = getenv("PARAKEET_B200_GEMM_TAIL_T2"); return (v ? atoi(v) : 600) / 1000.0; }();
    static const double t4 = [] { const char* v = getenv("PARAKEET_B200_GEMM_TAIL_T4"); return (v ? atoi(v) : 450) / 1000.0; }();
    int tail_c = 1, full_units = pair_tiles;
    if (tail_max >= 2 && g.map_w32 != nullptr && g.M_dev == nullptr && pair_tiles > pairs && pair_tiles % pairs != 0 && g.N % 256 == 0) {
      const int r = pair_tiles % pairs;
      double best = 1.0;
      for (int c = 2; c <= tail_max && c <= 4; c *= 2) {
        const double cost = (double)((r * c + pairs - 1) / pairs) * (c == 2 ? t2 : t4);
        if (cost < best - 1e-9) { best = cost; tail_c = c; }
      }
      if (tail_c > 1) full_units = pair_tiles - r;
    }
    const CUtensorMap& mw32 = tail_c > 1 ? *reinterpret_cast<const CUtensorMap*>(g.map_w32) : mw2;
    switch (g.epi.mode) {
#define PKB_GEMM_CASE2(MODE) case MODE: launch_cfg2<MODE>(pairs, ma2, mw2, mw32, g, lo_row_off2, tail_c, full_units, st); break;
      PKB_GEMM_CASE2(EPI_BIAS_F32) PKB_GEMM_CASE2(EPI_BIAS_RELU_F32) PKB_GEMM_CASE2(EPI_BIAS_RELU_ACT) PKB_GEMM_CASE2(EPI_BIAS_ROWMAP_F32)
      PKB_GEMM_CASE2(EPI_SILU_ACT) PKB_GEMM_CASE2(EPI_RESADD_F32) PKB_GEMM_CASE2(EPI_QKV) PKB_GEMM_CASE2(EPI_GLU_F32) PKB_GEMM_CASE2(EPI_F32)
      PKB_GEMM_CASE2(EPI_PARTIAL_F32)
#undef PKB_GEMM_CASE2
      default: PKB_CHECK(false, "gemm_tc: unknown epilogue mode");
    }
    return;
  }
  const int bn = g.batch > 1 ? 128 : g_force_bn == 128 || g_force_bn == 256 ? g_force_bn : pick_bn(g.M, g.N, sms);
  const int tiles = ((g.M + BM - 1) / BM) * ((g.N + bn - 1) / bn) * g.batch;
  const int grid = tiles < sms ? tiles : sms;
  const int lo_row_off = (int)(g.a_lo_off / g.lda);
  const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(&map_a);
  const CUtensorMap& mw = *reinterpret_cast<const CUtensorMap*>(&map_w);
  switch (g.epi.mode) {
#define PKB_GEMM_CASE(MODE)                                                                         \
    case MODE:                                                                                      \
      if (bn == 256) launch_cfg<256, MODE>(grid, ma, mw, g, lo_row_off, st);                        \
      else launch_cfg<128, MODE>(grid, ma, mw, g, lo_row_off, st);                                  \
      break;
    PKB_GEMM_CASE(EPI_BIAS_F32) PKB_GEMM_CASE(EPI_BIAS_RELU_F32) PKB_GEMM_CASE(EPI_BIAS_RELU_ACT) PKB_GEMM_CASE(EPI_BIAS_ROWMAP_F32)
    PKB_GEMM_CASE(EPI_SILU_ACT) PKB_GEMM_CASE(EPI_RESADD_F32) PKB_GEMM_CASE(EPI_QKV) PKB_GEMM_CASE(EPI_GLU_F32) PKB_GEMM_CASE(EPI_F32)
    PKB_GEMM_CASE(EPI_ACT)
#undef PKB_GEMM_CASE
    case EPI_PARTIAL_F32: {   // split-K: one work unit per (tile, split); 128-wide tiles, or 256-wide ones when the caller found enough units
      if (g.epi.part_wide && g.N % 256 == 0) {
        const int units = ((g.M + BM - 1) / BM) * (g.N / 256) * g.epi.splits;
        launch_cfg<256, EPI_PARTIAL_F32>(units < sms ? units : sms, ma, mw, g, lo_row_off, st);
      } else {
        const int units = ((g.M + BM - 1) / BM) * ((g.N + 127) / 128) * g.epi.splits;
        launch_cfg<128, EPI_PARTIAL_F32>(units < sms ? units : sms, ma, mw, g, lo_row_off, st);
      }
      break;
    }
    case EPI_ARGMAX: {      // slab geometry (kArgmaxParts) is defined for 256-wide tiles
      const int tiles256 = ((g.M + BM - 1) / BM) * ((g.N + 255) / 256);
      launch_cfg<256, EPI_ARGMAX>(tiles256 < sms ? tiles256 : sms, ma, mw, g, lo_row_off, st);
      break;
    }
    default: PKB_CHECK(false, "gemm_tc: unknown epilogue mode");
  }
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
