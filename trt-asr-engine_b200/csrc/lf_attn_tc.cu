// lf_attn_tc.cu -- whole-utterance relative-position attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Same arithmetic as lf_attention_tma_kernel (offline_long.cu; NeMo RelPositionMultiHeadAttention with full context):
//   scores[i][j] = ((q_i + u) . k_j + (q_i + v) . p_{i-j}) / sqrt(128),  softmax over j,  times V
// re-tiled for UMMA.  One CTA = 128 query rows of one (utterance, head); it walks the utterance in tiles of 64 keys:
//   S_t   = Qu K_t^T          UMMA M128 N64  K128   (content scores, TMEM)
//   G_b   = Qv Pblk_b^T       UMMA M128 N64  K128   (position scores of ONE new block of 64 relative positions per tile)
//   O    += P_t V_t           UMMA M128 N128 K64    (P_t: bf16 probabilities written to shared memory by the softmax warps)
// The [T, 2T-1] position-score matrix is never formed.  A 128 x 64 score tile needs relative positions
// D-63 .. D+127 (D = i0 - j0): three blocks of 64.  Two of them were already computed for the previous tile, so the blocks live in a
// 4-slot ring in TMEM and every tile computes exactly one new block: tensor work per tile = content + position + value = the
// algorithmic 3 x 128 x 64 x 128 MACs, nothing redundant (the mma.sync kernel recomputes a 127-row window per 64 x 64 tile).
// rel_shift is an index skew (score (r, c) uses window column r - c + 63).  In TMEM a thread owns one accumulator ROW and
// tcgen05.ld can only address columns uniformly across a warp, so the skew is resolved by letting lane l of a warp visit the keys
// in the rotated order c = (s + l) mod 64: then the window column r - c + 63 = 32 q + 63 - s (+ 64 after the wrap) IS uniform across
// the warp and the position scores come straight out of TMEM; the content score S[r][c] is the one read at a per-lane index, from
// the thread's own row of a shared-memory copy of S (written and read by the same thread: no synchronisation).
//
// TMEM (512 columns): O [0,128) | S double buffer [128,256) | position-block ring 4 x 64 [256,512).
// Shared memory: Qu, Qv (2 x 32 KB, resident), 2 stages x (K_t 16 KB + Pblk_{t+1} 16 KB + V^T_t 16 KB), P_t 16 KB, S copy 34 KB.
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocation), 4..11 = softmax: thread == (query row, half of the tile's 64 rotated
// key steps); warp % 4 == TMEM lane quarter.  The two threads of a row exchange their partial row maximum once per tile (named
// barrier of their two warps) and their partial row sums once at the end.  (One softmax warp per scheduler was the bottleneck of the
// first version: 472 TFLOP/s algorithmic at T = 45 000; see DESIGN.md.)
// V is consumed K-major, i.e. transposed ([d][key]): lf_prep_kernel writes V^T (and the two biased query planes) once per layer.
#include <cuda.h>

#include "offline_long.cuh"

namespace pkb {

namespace {

constexpr int kQT = 128;                       // query rows per CTA (UMMA M)
constexpr int kKT = 64;                        // keys per tile
constexpr int kSubA = 128 * 128;               // bytes of one [128 rows][64 bf16] 128-byte-swizzled sub-tile
constexpr int kSubB = 64 * 128;                // bytes of one [64 rows][64 bf16] sub-tile
constexpr int kOffQu = 0, kOffQv = 2 * kSubA;
constexpr int kOffStage = 4 * kSubA;           // 65536
constexpr int kStageK = 0, kStageP = 2 * kSubB, kStageV = 4 * kSubB, kStageBytes = 4 * kSubB + kSubA;      // 49152
constexpr int kOffP = kOffStage + 2 * kStageBytes;      // 163840: probabilities [128][64] bf16 (prologue: table block -1)
constexpr int kOffScr = kOffP + kSubA;                  // 180224: f32 copy of S, [128][kScrPitch] (prologue: table block 0)
constexpr int kScrPitch = 68;                           // floats per row: 16-byte aligned rows, conflict-free rotated reads (5 l + s mod 32)
constexpr int kOffBar = kOffScr + kQT * kScrPitch * 4;  // 215040
constexpr int kOffXch = kOffBar + 256;                  // float [2][128]: row maxima / row sums exchanged between the two column halves
constexpr size_t kSmemTc = 1024 + kOffXch + 4 * kQT * 4;      // [0,2): maxima, [2,4): final row sums
constexpr uint32_t kColO = 0, kColS = 128, kColG = 256;
constexpr float kScale2 = 0.08838834764831845f * 1.4426950408889634f;      // log2(e) / sqrt(128): scores are kept in the exp2 domain
constexpr float kTau2 = 8.0f;                           // O is rescaled only when a row maximum grows by more than this (factor 256)

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
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major operand tile, 128-byte swizzle, 8-row groups 1024 B apart (same descriptor as gemm_tc.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int N>
struct IDesc {      // kind::f16: D = f32, A = B = bf16, both K-major, M = 128
  static constexpr uint32_t value = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
};
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// two 32-column loads in flight behind ONE wait (each tcgen05.ld + wait pair exposes the full TMEM read latency)
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr_a, uint32_t taddr_b, uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]), "=r"(a[31]),
        "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]), "=r"(b[16]), "=r"(b[17]), "=r"(b[18]), "=r"(b[19]), "=r"(b[20]), "=r"(b[21]), "=r"(b[22]), "=r"(b[23]), "=r"(b[24]), "=r"(b[25]), "=r"(b[26]), "=r"(b[27]), "=r"(b[28]), "=r"(b[29]), "=r"(b[30]), "=r"(b[31])
      : "r"(taddr_a), "r"(taddr_b)
      : "memory");
}
// polling with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes (or the hint expires) instead
// of spinning through issue slots that the softmax warps on the same scheduler need
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  if (hint_ns == 0) { mbar_wait(bar, parity); return; }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\t"
      "tcgen05.wait::st.sync.aligned;" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// first V^T column of utterance e: the utterances sit at 64-aligned column offsets (whole key tiles never straddle two utterances)
__device__ __forceinline__ int vt_offset(const BatchDev& b, int e) {
  int off = 0;
  for (int x = 0; x < e; ++x) off += (b.Tq[x] + (kKT - 1)) & ~(kKT - 1);
  return off;
}

// 8 k-steps of a K = 128 product whose A and B tiles are stored as two 64-column swizzled sub-tiles each
__device__ __forceinline__ void issue_k128(uint32_t tmem_d, uint32_t a_addr, uint32_t a_sub_bytes, uint32_t b_addr, uint32_t b_sub_bytes,
                                           uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t da = make_desc(a_addr + (k >> 2) * a_sub_bytes) + (uint64_t)(2 * (k & 3));
    const uint64_t db = make_desc(b_addr + (k >> 2) * b_sub_bytes) + (uint64_t)(2 * (k & 3));
    umma(tmem_d, da, db, idesc, k != 0 ? 1u : 0u);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ per-layer operand preparation
// blocks [0, n_q):        Qu = q + pos_bias_u, Qv = q + pos_bias_v as bf16 planes [Mq][1024] (plane 1 starts q_plane elements later)
// blocks [n_q, n_q + n_v): V^T[h*128 + d][vt_offset(e) + j] = v_j[h*128 + d]   (lanes run along the key index: coalesced writes)
__global__ void __launch_bounds__(256)
lf_prep_kernel(BatchDev b, LfTcArgs a, int n_q_blocks) {
  pdl_enter();
  if ((int)blockIdx.x < n_q_blocks) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const int m = (int)(idx >> 7), c8 = (int)(idx & 127) * 8;
    if (m >= b.M) return;
    const uint4 raw = *reinterpret_cast<const uint4*>(a.qkv + (size_t)m * (3 * kDModel) + c8);
    const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
    uint32_t ou[4], ov[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 q = __bfloat1622float2(q2[k]);
      ou[k] = pack_bf16x2(q.x + a.bias_u[c8 + 2 * k], q.y + a.bias_u[c8 + 2 * k + 1]);
      ov[k] = pack_bf16x2(q.x + a.bias_v[c8 + 2 * k], q.y + a.bias_v[c8 + 2 * k + 1]);
    }
    *reinterpret_cast<uint4*>(a.q_planes + (size_t)m * kDModel + c8) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
    *reinterpret_cast<uint4*>(a.q_planes + a.q_plane + (size_t)m * kDModel + c8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    return;
  }
  // V^T: thread == (packed row m, 8-dim chunk); lane index runs along m
  const long long idx = (long long)(blockIdx.x - n_q_blocks) * 256 + threadIdx.x;
  const int m_round = (b.M + 255) & ~255;
  const int c8 = (int)(idx / m_round) * 8, m = (int)(idx % m_round);
  if (m >= b.M || c8 >= kDModel) return;
  const int e = b.row_entry[m];
  const long long col = vt_offset(b, e) + b.row_pos[m];
  const uint4 raw = *reinterpret_cast<const uint4*>(a.qkv + (size_t)m * (3 * kDModel) + 2 * kDModel + c8);
  const __nv_bfloat16* v8 = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
  for (int x = 0; x < 8; ++x) a.vt[(size_t)(c8 + x) * a.ldv + col] = v8[x];
}

// ------------------------------------------------------------------------------------------------ attention
// NH = softmax warpgroups per CTA: 1 (thread == query row, all 64 rotated key steps of a tile) or 2 (two threads per row, 32 steps each)
template <int NH>
__global__ void __launch_bounds__(128 + 128 * NH, 1)
lf_attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_pos, const __grid_constant__ CUtensorMap map_vt, BatchDev b, LfTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (a pointer rebuilt from an integer would be a generic pointer: every access through it became
  // LD.E / ST.E with 64-bit address arithmetic in the first version)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kOffBar);
  uint64_t* q_full = bars;            // prologue operands landed
  uint64_t* kp_full = bars + 1;       // [2] stage t: K_t and table block t+1 landed
  uint64_t* kp_empty = bars + 3;      // [2] S_t and position block t+1 have been computed: their operands are free (long before PV_t)
  uint64_t* sg_full = bars + 5;       // [2] S_t (buffer t & 1) and position block t+1 are in TMEM
  uint64_t* s_empty = bars + 7;       // [2] the softmax warps have copied S out of buffer t & 1
  uint64_t* p_full = bars + 9;        // P_t is in shared memory (and O has been rescaled)
  uint64_t* pv_done = bars + 10;      // PV_t has completed: O is readable, the P buffer is free
  uint64_t* v_full = bars + 11;       // [2] V^T_t landed
  uint64_t* v_empty = bars + 13;      // [2] PV_t has read V^T_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_pos) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vt) : "memory");
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kp_full[s], 1); mbar_init(&kp_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&sg_full[s], 1); mbar_init(&s_empty[s], 4 * NH);
    }
    mbar_init(p_full, 4 * NH);
    mbar_init(pv_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  pdl_enter();
  const int e = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kQT;
  const int T = b.Tq[e];
  const int n_tiles = i0 < T ? (T + kKT - 1) / kKT : 0;      // (all threads take the same path: the teardown below is common)
  const int row0 = b.row_off[e];

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && n_tiles > 0) {
      const int col_h = h * kDHead;
      const int voff = vt_offset(b, e);
      mbar_expect_tx(q_full, 4 * kSubA + 4 * kSubB);
      tma_load_2d(base + kOffQu, &map_q, q_full, col_h, row0 + i0);
      tma_load_2d(base + kOffQu + kSubA, &map_q, q_full, col_h + 64, row0 + i0);
      tma_load_2d(base + kOffQv, &map_q, q_full, col_h, a.q_plane_rows + row0 + i0);
      tma_load_2d(base + kOffQv + kSubA, &map_q, q_full, col_h + 64, a.q_plane_rows + row0 + i0);
      // table block beta holds relative positions i0 - 64 beta + 1 .. i0 - 64 beta + 64; table row = rel + (Tm - 1)
      const int prow_m1 = i0 + 64 + a.Tm, prow_0 = i0 + a.Tm;
      tma_load_2d(base + kOffP, &map_pos, q_full, col_h, prow_m1);
      tma_load_2d(base + kOffP + kSubB, &map_pos, q_full, col_h + 64, prow_m1);
      tma_load_2d(base + kOffScr, &map_pos, q_full, col_h, prow_0);
      tma_load_2d(base + kOffScr + kSubB, &map_pos, q_full, col_h + 64, prow_0);
      // K / table-block operands are released as soon as S_t and its position block have been computed, V^T_t only when PV_t has
      // run: two rings, so that the loads of tile t+2 start ~1.5 tiles before they are needed instead of right behind PV_t
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t & 1;
        uint8_t* st = base + kOffStage + s * kStageBytes;
        const uint32_t ph = ((t >> 1) & 1) ^ 1;
        mbar_wait_hint(&kp_empty[s], ph, a.wait_hint_ns);
        mbar_expect_tx(&kp_full[s], 4 * kSubB);
        tma_load_2d(st + kStageK, &map_k, &kp_full[s], kDModel + col_h, row0 + kKT * t);
        tma_load_2d(st + kStageK + kSubB, &map_k, &kp_full[s], kDModel + col_h + 64, row0 + kKT * t);
        const int prow = i0 - kKT * (t + 1) + a.Tm;      // block t + 1
        tma_load_2d(st + kStageP, &map_pos, &kp_full[s], col_h, prow);
        tma_load_2d(st + kStageP + kSubB, &map_pos, &kp_full[s], col_h + 64, prow);
        mbar_wait_hint(&v_empty[s], ph, a.wait_hint_ns);
        mbar_expect_tx(&v_full[s], kSubA);
        tma_load_2d(st + kStageV, &map_vt, &v_full[s], voff + kKT * t, col_h);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && n_tiles > 0) {
      const uint32_t uQu = smem_u32(base + kOffQu), uQv = smem_u32(base + kOffQv), uP = smem_u32(base + kOffP), uScr = smem_u32(base + kOffScr);
      const uint32_t uSt = smem_u32(base + kOffStage);
      mbar_wait(q_full, 0);
      mbar_wait(&kp_full[0], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // prologue: position blocks -1, 0 (ring slots 0, 1), S_0, position block 1 (slot 2)
      issue_k128(tmem_base + kColG + 0 * 64, uQv, kSubA, uP, kSubB, IDesc<64>::value);
      issue_k128(tmem_base + kColG + 1 * 64, uQv, kSubA, uScr, kSubB, IDesc<64>::value);
      issue_k128(tmem_base + kColS, uQu, kSubA, uSt + kStageK, kSubB, IDesc<64>::value);
      issue_k128(tmem_base + kColG + 2 * 64, uQv, kSubA, uSt + kStageP, kSubB, IDesc<64>::value);
      umma_commit(&sg_full[0]);
      umma_commit(&kp_empty[0]);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) {
          const int s1 = (t + 1) & 1;
          const uint32_t st1 = uSt + s1 * kStageBytes;
          mbar_wait_hint(&kp_full[s1], ((t + 1) >> 1) & 1, a.wait_hint_ns);
          mbar_wait_hint(&s_empty[s1], (((t + 1) >> 1) & 1) ^ 1, a.wait_hint_ns);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_k128(tmem_base + kColS + s1 * 64, uQu, kSubA, st1 + kStageK, kSubB, IDesc<64>::value);                 // S_{t+1}
          issue_k128(tmem_base + kColG + ((t + 3) & 3) * 64, uQv, kSubA, st1 + kStageP, kSubB, IDesc<64>::value);      // block t + 2
          umma_commit(&sg_full[s1]);
          umma_commit(&kp_empty[s1]);
        }
        mbar_wait_hint(&v_full[t & 1], (t >> 1) & 1, a.wait_hint_ns);
        mbar_wait_hint(p_full, t & 1, a.wait_hint_ns);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t vt = uSt + (t & 1) * kStageBytes + kStageV;
#pragma unroll
        for (int k = 0; k < 4; ++k)      // O += P_t V_t : K = 64 keys
          umma(tmem_base + kColO, make_desc(uP) + (uint64_t)(2 * k), make_desc(vt) + (uint64_t)(2 * k), IDesc<128>::value, (t | k) != 0 ? 1u : 0u);
        umma_commit(&v_empty[t & 1]);
        umma_commit(pv_done);
      }
    }
  } else if (warp >= 4 && n_tiles > 0) {
    // ===================================================== softmax: thread == (query row r of the tile, 64 / NH of the rotated key steps)
    constexpr int NS = 64 / NH;                                  // steps per thread
    constexpr int OC = 128 / NH;                                 // O columns this thread rescales / stores
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2, r = 32 * q + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    float* scr = reinterpret_cast<float*>(base + kOffScr) + r * kScrPitch;                  // this row of the f32 copy of S
    uint8_t* p_row = base + kOffP + r * 128;                                                // this row of the probability tile
    float* xch = reinterpret_cast<float*>(base + kOffXch);
    const uint32_t lane2 = 2u * lane, swz = (uint32_t)(r & 7) << 4;
    const int s0 = NS * half;                                    // this thread's steps: s0 .. s0 + NS - 1
    float m_run = -INFINITY, l_run = 0.f;                        // exp2 domain; l_run: this thread's partial row sum
#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
      const int bsel = t & 1, j0 = kKT * t;
      mbar_wait(&sg_full[bsel], (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- S row -> the row's shared-memory copy, so that it can be read at a per-lane index
      {
        if constexpr (NH == 1) {
          uint32_t v[32], w[32];
          tmem_ld32x2(lane_addr + kColS + bsel * 64, lane_addr + kColS + bsel * 64 + 32, v, w);
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            *reinterpret_cast<uint4*>(scr + 4 * x) = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
            *reinterpret_cast<uint4*>(scr + 32 + 4 * x) = make_uint4(w[4 * x], w[4 * x + 1], w[4 * x + 2], w[4 * x + 3]);
          }
        } else {
          uint32_t v[32];
          tmem_ld32(lane_addr + kColS + bsel * 64 + s0, v);
#pragma unroll
          for (int x = 0; x < 8; ++x)
            *reinterpret_cast<uint4*>(scr + s0 + 4 * x) = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if constexpr (NH == 2) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");      // both halves of the row copy are in place
        else __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[bsel]);
      }
      // ---- scores in the rotated key order c = (s + lane) mod 64; window column = 32 q + 63 - s (+ 64 once s + lane wraps)
      // window column w lives in ring slot (t + 2 - (w >> 6)) & 3 at column w & 63:  w < 64: block t+1,  < 128: block t,  else block t-1
      float x[NS];
      float mt = -INFINITY;
      const bool full_tile = j0 + kKT <= T;            // warp-uniform: only the last tile of an utterance masks keys
#pragma unroll
      for (int cc = 0; cc < NS / 32; ++cc) {
        const int sc0 = s0 + 32 * cc;                  // 0 or 32 (compile-time for NH == 1, warp-uniform for NH == 2)
        const int wa = 32 * q + 32 - sc0;              // first window column of the 32 loaded for the un-wrapped lanes
        uint32_t ga[32], gb[32];
        const bool wraps = sc0 == 32;                  // lanes with s + lane >= 64 exist only for s >= 33
        if (wraps) {
          const int wb = wa + 64;
          tmem_ld32x2(lane_addr + kColG + (((t + 2 - (wa >> 6)) & 3) << 6) + (wa & 63),
                      lane_addr + kColG + (((t + 2 - (wb >> 6)) & 3) << 6) + (wb & 63), ga, gb);
        } else {
          tmem_ld32(lane_addr + kColG + (((t + 2 - (wa >> 6)) & 3) << 6) + (wa & 63), ga);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int cw = lane + sc0 + j;                                   // s + lane
          float g = __uint_as_float(ga[31 - j]);
          if (wraps) g = cw >= 64 ? __uint_as_float(gb[31 - j]) : g;
          // (steps below 32 cannot wrap: base + immediate offset, no address arithmetic per element)
          const float sc = wraps ? scr[cw & 63] : scr[cw];
          const float v = (sc + g) * kScale2;
          x[32 * cc + j] = v;
          mt = fmaxf(mt, v);
        }
      }
      if (!full_tile) {                                // keys past the end of the utterance (last tile only)
        mt = -INFINITY;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const int c = (lane + s0 + j) & 63;
          x[j] = j0 + c < T ? x[j] : -INFINITY;
          mt = fmaxf(mt, x[j]);
        }
      }
      if constexpr (NH == 2) {                         // row maximum over both halves
        xch[half * kQT + r] = mt;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        mt = fmaxf(mt, xch[(half ^ 1) * kQT + r]);
      }
      // ---- online softmax with a lazy rescale of O (the running maximum only moves when it grows by more than kTau2)
      float alpha = 1.f;
      const bool grow = mt > m_run + kTau2;            // first tile: m_run = -inf -> true (key j0 is always valid: mt is finite)
      if (grow) { alpha = ex2f(m_run - mt); m_run = mt; }
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};             // four partial sums: no 64-deep dependent add chain
#pragma unroll
      for (int j = 0; j < NS; ++j) { x[j] = ex2f(x[j] - m_run); rs4[j & 3] += x[j]; }
      l_run = l_run * alpha + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      if (t > 0) {
        mbar_wait(pv_done, (t - 1) & 1);               // PV_{t-1} complete: the P buffer is free and O is stable
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
          for (int cc = 0; cc < OC / 32; ++cc) {
            uint32_t o[32];
            tmem_ld32(lane_addr + kColO + OC * half + 32 * cc, o);
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st32(lane_addr + kColO + OC * half + 32 * cc, o);
          }
        }
      }
      // ---- P_t -> shared memory, bf16, K-major 128-byte-swizzled A tile [128 rows][64 keys]: byte (2 c) ^ ((r & 7) << 4) of the row
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const uint32_t c2 = (lane2 + 2u * (uint32_t)(s0 + j)) & 127u;
        *reinterpret_cast<__nv_bfloat16*>(p_row + (c2 ^ swz)) = __float2bfloat16_rn(x[j]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's smem reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    if constexpr (NH == 2) {                           // row sums of the two halves
      xch[(2 + half) * kQT + r] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      l_run += xch[(2 + (half ^ 1)) * kQT + r];
    }
    // ---- context rows -> bf16 operand of linear_out
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const bool ok = i0 + r < T;
    __nv_bfloat16* dst = a.ctx + (size_t)(row0 + i0 + r) * a.ldc + h * kDHead + OC * half;
#pragma unroll 1
    for (int cc = 0; cc < OC / 32; ++cc) {
      uint32_t o[32];
      tmem_ld32(lane_addr + kColO + OC * half + 32 * cc, o);
      if (ok) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t w[4];
#pragma unroll
          for (int y = 0; y < 4; ++y) w[y] = pack_bf16x2(__uint_as_float(o[8 * k + 2 * y]) * inv, __uint_as_float(o[8 * k + 2 * y + 1]) * inv);
          *reinterpret_cast<uint4*>(dst + 32 * cc + 8 * k) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

void launch_lf_prep(const BatchDev& b, const LfTcArgs& a, cudaStream_t st) {
  if (b.M <= 0) return;
  const int n_q = (int)(((long long)b.M * 128 + 255) / 256);
  const long long m_round = (b.M + 255) & ~255;
  const int n_v = (int)((m_round * 128 + 255) / 256);
  launch_k(lf_prep_kernel, dim3(n_q + n_v), dim3(256), 0, st, b, a, n_q);
  PKB_CUDA(cudaGetLastError());
}

void launch_lf_attention_tc(const BatchDev& b, const LfTcArgs& a, int max_T, cudaStream_t st) {
  if (b.B <= 0 || max_T <= 0) return;
  static bool attr = false;
  if (!attr) {
    PKB_CUDA(cudaFuncSetAttribute(lf_attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTc));
    PKB_CUDA(cudaFuncSetAttribute(lf_attention_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTc));
    attr = true;
  }
  // PARAKEET_B200_LF_SPLIT: softmax warpgroups per CTA (1 or 2; A/B measurement knob)
  static const int nh = [] { const char* v = getenv("PARAKEET_B200_LF_SPLIT"); return v ? atoi(v) : 1; }();
  static const int hint = [] { const char* v = getenv("PARAKEET_B200_LF_HINT_NS"); return v ? atoi(v) : 0; }();
  LfTcArgs a2 = a;
  a2.wait_hint_ns = hint;
  const dim3 grid((max_T + kQT - 1) / kQT, kHeads, b.B);
  const CUtensorMap& mq = *reinterpret_cast<const CUtensorMap*>(a.map_q);
  const CUtensorMap& mk = *reinterpret_cast<const CUtensorMap*>(a.map_k);
  const CUtensorMap& mp = *reinterpret_cast<const CUtensorMap*>(a.map_pos);
  const CUtensorMap& mv = *reinterpret_cast<const CUtensorMap*>(a.map_vt);
  if (nh == 2) launch_k(lf_attention_tc_kernel<2>, grid, dim3(384), kSmemTc, st, mq, mk, mp, mv, b, a2);
  else launch_k(lf_attention_tc_kernel<1>, grid, dim3(256), kSmemTc, st, mq, mk, mp, mv, b, a2);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
