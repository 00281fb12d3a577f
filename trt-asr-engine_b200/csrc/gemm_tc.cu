// gemm_tc.cu -- tcgen05 / TMA / TMEM GEMM for sm_100a (see gemm.h):  C[m,n] = sum_k A[m,k] * W[n,k], fused epilogue.
//
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d of a 128x64 bf16 A tile and a 128x64 bf16 W tile per k-block
//              into a kStages-deep shared-memory ring (128-byte swizzle), completion on "full" mbarriers
//   warp 1   : MMA issuer    -- one elected lane issues 4 x tcgen05.mma (M128 N128 K16, bf16 -> f32) per k-block with the
//              accumulator in TMEM (128 lanes x 128 columns); tcgen05.commit releases the smem slot ("empty" mbarrier)
//              and finally signals the epilogue ("tmem_full" mbarrier).  Warp 1 also owns the TMEM allocation.
//   warps 2-5: epilogue      -- tcgen05.ld 32 lanes x 16 columns at a time (warp w reads TMEM lane quarter w%4),
//              apply the fused epilogue (bias / ReLU / SiLU / GLU / residual add / QKV ring scatter) and store.
// Split ("precise") mode runs the k-loop twice over the same W tiles: first the bf16 high plane of A, then the low
// plane, accumulating into the same TMEM tile -- fp32-grade products at 2x the tensor work, no extra weight traffic
// from HBM (the W tile of the second pass hits L2).
#include <cuda.h>

#include "gemm.h"

namespace pkb {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kStages = 6;
constexpr int kTileBytes = BM * BK * 2;          // 16 KB (A and W tiles have the same size)
constexpr int kThreads = 192;
constexpr int kTmemCols = 128;
constexpr size_t kSmemBytes = 1024 + (size_t)kStages * 2 * kTileBytes + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
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

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const GemmArgs g,
               const int lo_row_off) {
  extern __shared__ uint8_t smem_raw[];
  const int M = g.M_dev ? min(*g.M_dev, g.M) : g.M;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M) return;                                   // uniform per CTA, before any barrier / allocation

  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + kStages * kTileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + 2 * kStages * kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_per_pass = g.K / BK;
  const int num_kb = kb_per_pass * (g.a_lo_off != 0 ? 2 : 1);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], 2 * kTileBytes);
        const int pass = kb / kb_per_pass, kk = (kb % kb_per_pass) * BK;
        tma_load_2d(sA + s * kTileBytes, &map_a, &full_bar[s], kk, m0 + pass * lo_row_off);
        tma_load_2d(sB + s * kTileBytes, &map_w, &full_bar[s], kk, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t da = make_smem_desc(smem_u32(sA + s * kTileBytes));
        const uint64_t db = make_smem_desc(smem_u32(sB + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)                  // +32 bytes per K=16 step inside the 128-byte swizzle atom
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);                       // smem slot reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);                         // accumulator complete
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quarters 2,3,0,1
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const int n = n0 + c0 + j;
          if (n < g.N) epilogue_pair(g.epi, m, n, g.N, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
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

bool gemm_tc_supported(const GemmArgs& g) {
  return g.K % BK == 0 && g.lda == g.K && (g.a_lo_off % g.lda) == 0 && g.M > 0 && g.N > 0;
}

void gemm_tc(const GemmArgs& g, const TensorMap& map_a, const TensorMap& map_w, cudaStream_t st) {
  PKB_CHECK(gemm_tc_supported(g), "gemm_tc: unsupported shape");
  static bool attr = false;
  if (!attr) {
    PKB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr = true;
  }
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  const int lo_row_off = (int)(g.a_lo_off / g.lda);
  gemm_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(*reinterpret_cast<const CUtensorMap*>(&map_a),
                                                     *reinterpret_cast<const CUtensorMap*>(&map_w), g, lo_row_off);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
