// attn_mma.cu -- relative-position attention over [ring cache (256) || new rows (Tq)], bf16 tensor-core version.
//
// Semantics: NeMo RelPositionMultiHeadAttention with a cache (see enc_kernels.cu); scores (q+u)K^T + rel_shift((q+v)P^T),
// scaled by 1/sqrt(128), masked to the valid cache suffix and the new rows, softmax, times V.
//
// Persistent kernel, one CTA per SM, work item = (stream, head).  HBM-bound by design: per item it must read the 288 x 128
// bf16 K and V ring slices (147 KB), everything else is on-chip.
//   * K and V rings are head-major [slot][head][288][128] bf16, so one (stream, head) block of 96 keys is 24 KB of
//     CONTIGUOUS HBM (a [288][1024] layout would read it as 128-byte pieces at a 2 KB stride, which caps HBM efficiency
//     near 60 %); the slices are fetched with TMA (96-key x 64-dim boxes, 128-byte swizzle)
//     by a dedicated producer warp into a 6-stage shared-memory ring (full/empty mbarriers): the producer runs a whole
//     item ahead of the 8 consumer warps, so the loads of item i+1 overlap the math of item i.
//   * only VALID ring slots are fetched: a 96-key block whose 12 groups of 8 keys are all valid is two 96-row boxes; a block that
//     contains invalid slots (the 26 slots of the 288-slot ring outside [cache suffix || new rows] in steady state, most of the
//     ring while a stream's cache is still filling) is fetched as its valid groups only, through 32-row / 8-row boxes of the same
//     rings at the same shared-memory positions.  Rows that were not fetched hold stale data: their scores are masked by select
//     and their V fragments are zeroed in registers, so nothing stale is ever consumed arithmetically.  (Round 1 fetched whole
//     blocks: ncu dram__bytes_read 1.252 GB per launch against 1.099 GB of valid K / V rows at 1024 saturated streams.)
//   * keys are walked in PHYSICAL ring order (the softmax sum does not care), so a wrapped FIFO needs no second copy and
//     every box is one contiguous row range; validity and the relative position of slot p follow from the ring head.
//   * the position term G = (q+v) P^T is NOT computed here: it is the same P for every stream, so one batched tcgen05 GEMM
//     per layer produces it for all streams and heads (engine.cu) and this kernel only stages its rows;
//     S = (q+u) K^T (B fragments via ldmatrix from the swizzled tiles) and O = P V (ldmatrix.trans on the V tiles) run
//     on mma.sync m16n8k16 (bf16 -> f32), softmax in fp32 in smem.  R = query rows handled (8: the steady-state 6-row
//     chunk, 16, 32).
#include <cuda.h>

#include "enc_kernels.cuh"

namespace pkb {

namespace {

constexpr int kBlkKeys = 96;                       // keys per TMA block; kRingCap = 3 blocks
constexpr int kNumBlk = kRingCap / kBlkKeys;
constexpr int kBoxBytes = kBlkKeys * 128;          // one 96 x 64 bf16 box
constexpr int kStageBytes = 2 * kBoxBytes;         // d 0..63 and d 64..127
constexpr int kSPitch = 296;                       // floats; 296 % 32 == 8 -> conflict-free float2 fragment stores
constexpr int kGPitch = 328;                       // bf16 elements per staged row of position scores (covers the 320 table rows; 16-byte multiple)
constexpr int kPPitchB = 592;                      // bytes per probability row (296 bf16); 37 x 16 B -> conflict-free ldmatrix
static_assert(kRingCap % kBlkKeys == 0 && kPosRowsPad >= kPosRows, "ring / table geometry");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// K / V ring blocks are read exactly once per step: loading them with an L2 evict-first policy keeps the 1.2 GB-per-layer stream
// from flushing the residual stream, the GEMM operands and the partial sums (all re-read within the layer) out of the 126 MB L2
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
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

// Launch shape: CW consumer warps (+1 producer warp) per CTA, kStages ring stages, kCtasPerSm co-resident CTAs.
// PARAKEET_B200_ATTN_CFG selects among the compiled shapes for R = 8 (A/B measurements; default 0).

template <int R, int CFG>
struct Smem {
  // R = 8 (the steady-state chunk), CFG 0 (default): 4 consumer warps, 2 stages, THREE CTAs per SM -- the consumers of one item
  // run a chain of short dependent phases, so throughput comes from independent items in flight (measured 245 vs 316 us per
  // layer at 1024 streams against CFG 1: 8 consumer warps, 3 stages, two CTAs per SM).  Larger variants: one CTA per SM.
  static constexpr int kCW = (R == 8 && CFG == 0) ? 4 : 8;
  static constexpr int kStages = R == 8 ? (CFG == 0 ? 2 : 3) : R == 16 ? 5 : 4;
  static constexpr int kCtasPerSm = R == 8 ? (CFG == 0 ? 3 : 2) : 1;
  static constexpr int kThreads = 32 * (1 + kCW);
  static constexpr int kS = R * kSPitch * 4;
  static constexpr int kG = (R < 16 ? 8 : R) * (kGPitch * 2 > kPPitchB ? kGPitch * 2 : kPPitchB);   // bf16 G rows, later the bf16 probabilities
  static constexpr size_t kBytes = 1024 + (size_t)kStages * kStageBytes + kS + kG + 128;
};

constexpr int kGrpPerBlk = kBlkKeys / 8;          // 8-key groups per block
struct Item {
  int Tq, qlen, len, head, row0, ring_row0;
  unsigned need;      // bit k: 96-key block k holds at least one valid key
  unsigned long long groups;      // bit q: the 8-key group of physical slots [8q, 8q+8) holds at least one valid key
};
__device__ __forceinline__ Item load_item(const BatchDev& b, const AttnMmaArgs& a, int e) {
  Item it;
  it.Tq = b.Tq[e]; it.qlen = b.qlen[e]; it.len = b.len[e]; it.head = b.head[e]; it.row0 = b.row_off[e];
  it.ring_row0 = (a.layer * a.n_slots + b.slot[e]) * (kHeads * kRingCap);      // + head * kRingCap: rings are head-major
  // valid logical positions j in [256-len, 256+qlen) form one circular run of physical slots
  unsigned long long m = ring_valid_groups(it.head, it.len, it.qlen);      // common.cuh
  if (!a.trim) {      // whole blocks (A/B switch): every group of a needed block counts as valid
    unsigned long long full = 0;
#pragma unroll
    for (int k = 0; k < kNumBlk; ++k)
      if ((m >> (k * kGrpPerBlk)) & ((1ull << kGrpPerBlk) - 1ull)) full |= ((1ull << kGrpPerBlk) - 1ull) << (k * kGrpPerBlk);
    m = full;
  }
  it.groups = m;
  it.need = 0;
#pragma unroll
  for (int k = 0; k < kNumBlk; ++k)
    if ((m >> (k * kGrpPerBlk)) & ((1ull << kGrpPerBlk) - 1ull)) it.need |= 1u << k;
  return it;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int THREADS>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }

// Persistent: one CTA per SM walks (stream, head) items; the producer warp streams the K then V blocks of successive items
// through the ring without waiting for the math, so HBM stays busy while the consumers are in their on-chip phases.
// R = 8 (rows 8..15 of every m16 tile are identically zero and never loaded / stored), 16 or 32.
template <int R, int CFG>
__global__ void __launch_bounds__(Smem<R, CFG>::kThreads, Smem<R, CFG>::kCtasPerSm)
attention_mma_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                     const __grid_constant__ CUtensorMap map_k32, const __grid_constant__ CUtensorMap map_v32,
                     const __grid_constant__ CUtensorMap map_k8, const __grid_constant__ CUtensorMap map_v8, BatchDev b, AttnMmaArgs a) {
  constexpr int MT = R <= 16 ? 1 : 2;
  constexpr bool kHalf = R == 8;
  constexpr int kStages = Smem<R, CFG>::kStages;
  constexpr int kConsumerWarps = Smem<R, CFG>::kCW;
  constexpr int kConsumerThreads = 32 * kConsumerWarps;
  constexpr int NTW = 16 / kConsumerWarps;        // 8-wide head-dim tiles per consumer warp in the PV phase (2 or 4)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* s_ring = base;
  float* s_S = reinterpret_cast<float*>(base + kStages * kStageBytes);
  __nv_bfloat16* s_G = reinterpret_cast<__nv_bfloat16*>(s_S + R * kSPitch);      // [R][kGPitch] bf16, later the probabilities
  uint8_t* s_P = reinterpret_cast<uint8_t*>(s_G);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_G) + Smem<R, CFG>::kG);
  uint64_t* empty_bar = full_bar + kStages;

  const int n_items = b.B * kHeads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k32) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v32) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k8) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v8) : "memory");
#pragma unroll
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_enter();          // set-up above overlaps the previous kernel's tail; no global data is touched before this point
  __syncthreads();

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint64_t pol_stream = 0;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      const bool hint = a.evict_first != 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int e = item >> 3, h = item & 7;
        const Item im = load_item(b, a, e);
#pragma unroll
        for (int kv = 0; kv < 2; ++kv) {
          const CUtensorMap* map = kv ? &map_v : &map_k;
          const CUtensorMap* map32 = kv ? &map_v32 : &map_k32;
          const CUtensorMap* map8 = kv ? &map_v8 : &map_k8;
#pragma unroll
          for (int k = 0; k < kNumBlk; ++k) {
            if (!((im.need >> k) & 1u)) continue;
            const unsigned gm = (unsigned)(im.groups >> (k * kGrpPerBlk)) & ((1u << kGrpPerBlk) - 1u);
            const int s = it % kStages;
            mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
            uint8_t* dst = s_ring + s * kStageBytes;
            const int row = im.ring_row0 + h * kRingCap + k * kBlkKeys;
            if (gm == (1u << kGrpPerBlk) - 1u) {
              mbar_expect_tx(&full_bar[s], kStageBytes);
              if (hint) {
                tma_load_2d_hint(dst, map, &full_bar[s], 0, row, pol_stream);
                tma_load_2d_hint(dst + kBoxBytes, map, &full_bar[s], 64, row, pol_stream);
              } else {
                tma_load_2d(dst, map, &full_bar[s], 0, row);
                tma_load_2d(dst + kBoxBytes, map, &full_bar[s], 64, row);
              }
            } else {
              // valid groups only: aligned runs of four groups as one 32-row box, the rest as 8-row boxes; a group of 8 keys is
              // 8 rows x 128 B = one 1024-byte swizzle atom of each 64-dim half, so partial boxes land exactly where the 96-row
              // box would have put the same rows
              mbar_expect_tx(&full_bar[s], (uint32_t)__popc(gm) * 2048u);
#pragma unroll
              for (int q4 = 0; q4 < kGrpPerBlk / 4; ++q4) {
                const unsigned g4 = (gm >> (4 * q4)) & 15u;
                if (g4 == 15u) {
                  uint8_t* d = dst + q4 * 4096;
                  if (hint) {
                    tma_load_2d_hint(d, map32, &full_bar[s], 0, row + 32 * q4, pol_stream);
                    tma_load_2d_hint(d + kBoxBytes, map32, &full_bar[s], 64, row + 32 * q4, pol_stream);
                  } else {
                    tma_load_2d(d, map32, &full_bar[s], 0, row + 32 * q4);
                    tma_load_2d(d + kBoxBytes, map32, &full_bar[s], 64, row + 32 * q4);
                  }
                } else {
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    if (!((g4 >> q) & 1u)) continue;
                    uint8_t* d = dst + (4 * q4 + q) * 1024;
                    const int r8 = row + 8 * (4 * q4 + q);
                    if (hint) {
                      tma_load_2d_hint(d, map8, &full_bar[s], 0, r8, pol_stream);
                      tma_load_2d_hint(d + kBoxBytes, map8, &full_bar[s], 64, r8, pol_stream);
                    } else {
                      tma_load_2d(d, map8, &full_bar[s], 0, r8);
                      tma_load_2d(d + kBoxBytes, map8, &full_bar[s], 64, r8);
                    }
                  }
                }
              }
            }
            ++it;
          }
        }
      }
    }
    return;
  }

  // ===== consumers =====
  const int cw = warp - 1, g = lane >> 2, t = lane & 3;
  const float scale = 0.08838834764831845f;       // 1/sqrt(128)
  const uint32_t sp = smem_u32(s_P);
  int it = 0;
#pragma unroll 1
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int e = item >> 3, h = item & 7;
    const Item im = load_item(b, a, e);
    const int Tq = im.Tq;

    // ---- query fragments (q + pos_bias_u, bf16 plane written by the QKV epilogue; natural k order = ldmatrix order of the K tiles)
    uint32_t qu[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int i = 16 * mt + g + 8 * hr;
        const bool ok = i < Tq && !(kHalf && hr == 1);
        const __nv_bfloat16* qrow = a.q_bf16 + (size_t)(im.row0 + (ok ? i : 0)) * kDModel + h * kDHead;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          qu[mt][ks][hr] = ok ? *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * t) : 0u;
          qu[mt][ks][hr + 2] = ok ? *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 8 + 2 * t) : 0u;
        }
      }
    }

    // ---- position scores G[i][r] = (q_i + pos_bias_v) . P[r]: one batched tcgen05 GEMM per layer computes them for every
    //      stream and head (engine.cu); here the item's rows are staged in shared memory (bf16 [rows][kGPitch])
    {
      const int cthread = tid - 32;
      for (int x = cthread; x < Tq * (kPosRowsPad / 8); x += kConsumerThreads) {
        const int i = x / (kPosRowsPad / 8), c = x % (kPosRowsPad / 8);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.g_pos + (size_t)(im.row0 + i) * (kHeads * kPosRowsPad) + h * kPosRowsPad) + c);
        *reinterpret_cast<uint4*>(s_G + i * kGPitch + 8 * c) = v;
      }
    }
    consumer_sync<kConsumerThreads>();      // G complete

    // ---- S phase: scores for every needed key block, combined with the skewed position term, masked, scaled
#pragma unroll 1
    for (int k = 0; k < kNumBlk; ++k) {
      if (!((im.need >> k) & 1u)) continue;
      const int s = it % kStages;
      mbar_wait(&full_bar[s], (it / kStages) & 1);
      ++it;
      const uint32_t st = smem_u32(s_ring + s * kStageBytes);
#pragma unroll 1
      for (int nb = cw; nb < kBlkKeys / 8; nb += kConsumerWarps) {
        float c[MT][4], cb[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          c[mt][0] = c[mt][1] = c[mt][2] = c[mt][3] = 0.f;
          cb[mt][0] = cb[mt][1] = cb[mt][2] = cb[mt][3] = 0.f;
        }
        const int key_l = 8 * nb + (lane & 7), mi = lane >> 3;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int chunk = 4 * (u & 1) + mi;
          const uint32_t addr = st + (u >> 1) * kBoxBytes + key_l * 128 + ((chunk ^ (key_l & 7)) << 4);
          uint32_t r0, r1, r2, r3;
          ldsm_x4(addr, r0, r1, r2, r3);
          const int ks = 4 * (u >> 1) + 2 * (u & 1);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            mma_bf16(c[mt], qu[mt][ks], r0, r1);
            mma_bf16(cb[mt], qu[mt][ks + 1], r2, r3);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int x = 0; x < 4; ++x) c[mt][x] += cb[mt][x];
#pragma unroll
          for (int hr = 0; hr < (kHalf ? 1 : 2); ++hr) {
            const int i = 16 * mt + g + 8 * hr;
            if (i >= Tq) continue;
            float out[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              const int p = k * kBlkKeys + 8 * nb + 2 * t + x;
              int j = p - im.head;
              j += j < 0 ? kRingCap : 0;
              const bool valid = j >= kCacheS - im.len && j < kCacheS + im.qlen;
              out[x] = valid ? (c[mt][2 * hr + x] + __bfloat162float(s_G[i * kGPitch + (kCacheS + i - j) + kPosNeg])) * scale : -INFINITY;
            }
            *reinterpret_cast<float2*>(s_S + i * kSPitch + k * kBlkKeys + 8 * nb + 2 * t) = make_float2(out[0], out[1]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);      // this warp is done with the K block
    }
    consumer_sync<kConsumerThreads>();      // scores complete; nobody reads s_G any more

    // ---- softmax (fp32) -> bf16 probabilities in the A-operand layout [rows][296] (rows >= Tq and skipped blocks are zero)
    constexpr int RP = kHalf ? 8 : 16 * MT;
    for (int i = cw; i < RP; i += kConsumerWarps) {
      __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(s_P + i * kPPitchB);
      float v[kRingCap / 32];
      float mx = -INFINITY;
#pragma unroll
      for (int x = 0; x < kRingCap / 32; ++x) {
        const int p = lane + 32 * x;
        v[x] = (i < Tq && ((im.need >> (p / kBlkKeys)) & 1u)) ? s_S[i * kSPitch + p] : -INFINITY;
        mx = fmaxf(mx, v[x]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int x = 0; x < kRingCap / 32; ++x) {
        v[x] = mx > -INFINITY ? __expf(v[x] - mx) : 0.f;
        sum += v[x];
      }
      sum = warp_sum(sum);
      const float inv = (i < im.qlen && sum > 0.f) ? 1.f / sum : 0.f;      // padded query rows are fully masked -> zeros
#pragma unroll
      for (int x = 0; x < kRingCap / 32; ++x) prow[lane + 32 * x] = __float2bfloat16_rn(v[x] * inv);
    }
    consumer_sync<kConsumerThreads>();

    // ---- O = P V : consumer warp cw owns head dims [16 cw, 16 cw + 16)
    float o[MT][NTW][4], ob[MT][NTW][4];      // even / odd k-steps accumulate separately (shorter dependent-HMMA chains)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTW; ++nt) {
        o[mt][nt][0] = o[mt][nt][1] = o[mt][nt][2] = o[mt][nt][3] = 0.f;
        ob[mt][nt][0] = ob[mt][nt][1] = ob[mt][nt][2] = ob[mt][nt][3] = 0.f;
      }
#pragma unroll 1
    for (int k = 0; k < kNumBlk; ++k) {
      if (!((im.need >> k) & 1u)) continue;
      const int s = it % kStages;
      mbar_wait(&full_bar[s], (it / kStages) & 1);
      ++it;
      const uint32_t st = smem_u32(s_ring + s * kStageBytes);
      const unsigned gm = (unsigned)(im.groups >> (k * kGrpPerBlk)) & ((1u << kGrpPerBlk) - 1u);
#pragma unroll
      for (int kk = 0; kk < kBlkKeys / 16; ++kk) {
        const bool lo_ok = (gm >> (2 * kk)) & 1u, hi_ok = (gm >> (2 * kk + 1)) & 1u;      // keys 0..7 / 8..15 of this step were fetched
        if (!lo_ok && !hi_ok) continue;      // their probabilities are zero
        const int key0 = k * kBlkKeys + 16 * kk;
        uint32_t af[MT][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          if (kHalf) {
            const int l = lane & 15;      // .x2 takes its row addresses from lanes 0..15
            ldsm_x2(sp + (l & 7) * kPPitchB + (key0 + (l >> 3) * 8) * 2, af[mt][0], af[mt][2]);
            af[mt][1] = 0u; af[mt][3] = 0u;
          } else {
            const int mi = lane >> 3, ri = lane & 7;
            ldsm_x4(sp + (16 * mt + (mi & 1) * 8 + ri) * kPPitchB + (key0 + (mi >> 1) * 8) * 2, af[mt][0], af[mt][1], af[mt][2], af[mt][3]);
          }
        }
        const int mi = lane >> 3, ri = lane & 7;
        const int key_l = 16 * kk + (mi & 1) * 8 + ri;
#pragma unroll
        for (int np = 0; np < NTW / 2; ++np) {
          const int chunk_g = NTW * cw + 2 * np + (mi >> 1);     // 16-byte chunk over the 128 head dims
          const uint32_t addr = st + (chunk_g >> 3) * kBoxBytes + key_l * 128 + (((chunk_g & 7) ^ (key_l & 7)) << 4);
          uint32_t r0, r1, r2, r3;
          ldsm_x4_t(addr, r0, r1, r2, r3);
          // rows that were not fetched hold stale bytes (possibly NaN patterns): 0 x NaN must not reach the accumulators
          r0 = lo_ok ? r0 : 0u; r2 = lo_ok ? r2 : 0u;
          r1 = hi_ok ? r1 : 0u; r3 = hi_ok ? r3 : 0u;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (kk & 1) { mma_bf16(ob[mt][2 * np], af[mt], r0, r1); mma_bf16(ob[mt][2 * np + 1], af[mt], r2, r3); }
            else { mma_bf16(o[mt][2 * np], af[mt], r0, r1); mma_bf16(o[mt][2 * np + 1], af[mt], r2, r3); }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // ---- context rows -> bf16 operand of the output projection
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hr = 0; hr < (kHalf ? 1 : 2); ++hr) {
        const int i = 16 * mt + g + 8 * hr;
        if (i >= Tq) continue;
        __nv_bfloat16* dst = a.ctx.ptr + (size_t)(im.row0 + i) * a.ctx.lda + h * kDHead + 8 * NTW * cw + 2 * t;
#pragma unroll
        for (int nt = 0; nt < NTW; ++nt)
          *reinterpret_cast<uint32_t*>(dst + 8 * nt) =
              pack_bf16x2(o[mt][nt][2 * hr] + ob[mt][nt][2 * hr], o[mt][nt][2 * hr + 1] + ob[mt][nt][2 * hr + 1]);
      }
    consumer_sync<kConsumerThreads>();      // every warp is done with s_P before the next item's G phase overwrites it
  }
}

template <int R, int CFG>
void launch_r(const BatchDev& b, const AttnMmaArgs& a, int sms, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    PKB_CUDA(cudaFuncSetAttribute(attention_mma_kernel<R, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<R, CFG>::kBytes));
    attr = true;
  }
  const int items = b.B * kHeads;
  const int ctas = sms * Smem<R, CFG>::kCtasPerSm;
  launch_k(attention_mma_kernel<R, CFG>, dim3(items < ctas ? items : ctas), dim3(Smem<R, CFG>::kThreads), Smem<R, CFG>::kBytes, st,
           *reinterpret_cast<const CUtensorMap*>(a.map_k), *reinterpret_cast<const CUtensorMap*>(a.map_v),
           *reinterpret_cast<const CUtensorMap*>(a.map_k32), *reinterpret_cast<const CUtensorMap*>(a.map_v32),
           *reinterpret_cast<const CUtensorMap*>(a.map_k8), *reinterpret_cast<const CUtensorMap*>(a.map_v8), b, a);
}

}  // namespace

void launch_attention_mma(const BatchDev& b, const AttnMmaArgs& a, cudaStream_t st) {
  if (b.B <= 0) return;
  PKB_CHECK(a.ctx.lo_off == 0, "attention_mma: bf16 mode only");
  static const int evict = [] { const char* v = getenv("PARAKEET_B200_ATTN_EVICT"); return (v && v[0] == '0') ? 0 : 1; }();
  static const int trim = [] { const char* v = getenv("PARAKEET_B200_ATTN_TRIM"); return (v && v[0] == '0') ? 0 : 1; }();
  AttnMmaArgs a2 = a;
  a2.evict_first = evict;
  a2.trim = (trim && a.map_k32 && a.map_v32 && a.map_k8 && a.map_v8) ? 1 : 0;
  if (!a2.trim) { a2.map_k32 = a2.map_k8 = a.map_k; a2.map_v32 = a2.map_v8 = a.map_v; }      // never dereferenced by the kernel then
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    PKB_CUDA(cudaGetDevice(&dev));
    PKB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static int cfg = -1;
  if (cfg < 0) { const char* v = getenv("PARAKEET_B200_ATTN_CFG"); cfg = v ? atoi(v) : 0; }
  if (b.max_Tq <= 8) { if (cfg == 1) launch_r<8, 1>(b, a2, sms, st); else launch_r<8, 0>(b, a2, sms, st); }
  else if (b.max_Tq <= 16) launch_r<16, 0>(b, a2, sms, st);
  else launch_r<32, 0>(b, a2, sms, st);
}

}  // namespace pkb
