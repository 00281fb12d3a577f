// common.cuh -- shared helpers for the sm_100a kernels of libparakeet_trt (B200-native build).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

namespace pkb {

// ---- model constants (published Parakeet-TDT-0.6B-v3 architecture; contract.json:54-66,161-215) ----
constexpr int kNMels = 128;
constexpr int kDModel = 1024;
constexpr int kHeads = 8;
constexpr int kDHead = 128;
constexpr int kFF = 4096;
constexpr int kConvK = 9;
constexpr int kSubCh = 256;
constexpr int kCacheS = 256;     // last_channel_cache_size
constexpr int kRingCap = 288;    // physical ring capacity: 256 cached + up to 32 rows of the running chunk
constexpr int kTimeCtx = 4;      // conv time cache columns
constexpr int kCacheDrop = 3;
constexpr int kValidOut = 3;
constexpr int kDropPre = 2;
constexpr int kVocab = 8193;     // token head incl. blank
constexpr int kBlank = 8192;
constexpr int kNDur = 5;
constexpr int kJointOut = kVocab + kNDur;  // 8198
constexpr int kPredH = 640;
constexpr int kPredL = 2;
constexpr int kJointH = 640;
constexpr int kMaxSymbols = 8;
constexpr int kMaxTq = 32;       // T<=256 frames per push -> 32 tokens (streaming drops 2 of them, offline keeps all)
constexpr int kPosRows = kCacheS + 2 * kMaxTq;  // relative positions -(kMaxTq-1) .. 256+kMaxTq-1 (padded)
constexpr int kPosNeg = kMaxTq - 1;             // table row of relative position r is (r + kPosNeg)
constexpr int kPosRowsPad = 320;                // rows of the natural-layout table (multiple of 8, zero padded)

// Valid 8-slot groups of a K / V ring (attn_mma.cu fetches only these).  The valid logical positions of an entry -- the cache suffix
// [256 - len, 256) and the qlen new rows [256, 256 + qlen) -- are ONE circular run of physical slots starting at
// (head + 256 - len) mod kRingCap; bit q of the result is set when group q = slots [8q, 8q + 8) intersects that run.
// Host-callable so that the bit algebra is checked exhaustively on the CPU (pkb_debug_ring_valid_groups, tests/test_ring_groups.py).
#ifdef __CUDACC__
__host__ __device__
#endif
inline unsigned long long ring_valid_groups(int head, int len, int qlen) {
  constexpr int kGroups = kRingCap / 8;
  const int v_start = (head + kCacheS - len) % kRingCap, v_cnt = len + qlen;
  const int q0 = v_start >> 3;
  int n = ((v_start + (v_cnt > 0 ? v_cnt : 1) - 1) >> 3) - q0 + 1;      // groups q0 .. q0 + n - 1, circular
  n = n > kGroups ? kGroups : n;
  const int hi = q0 + n < kGroups ? q0 + n : kGroups;
  unsigned long long m = ((1ull << hi) - 1ull) & ~((1ull << q0) - 1ull);
  if (q0 + n > kGroups) m |= (1ull << (q0 + n - kGroups)) - 1ull;
  return m;
}

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define PKB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw ::pkb::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                             __FILE__ + ":" + std::to_string(__LINE__));                        \
  } while (0)

#define PKB_CHECK(cond, msg)                                                     \
  do {                                                                           \
    if (!(cond)) throw std::runtime_error(std::string("check failed: ") + (msg)); \
  } while (0)

// Programmatic dependent launch: every hot-path kernel is launched with the stream-serialisation attribute, lets its
// successor start scheduling immediately (pdl_trigger) and does its own set-up before pdl_wait(), which returns once the
// predecessor grid has completed and flushed.  No kernel touches global data before pdl_wait().
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PARAKEET_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
// While a CUDA-graph body is being captured the launch attribute can be withheld (graph_pdl_suppressed()): the kernels then depend
// on their predecessors through ordinary graph edges and their griddepcontrol instructions are no-ops.
inline bool& graph_pdl_suppressed() {
  static thread_local bool v = false;
  return v;
}
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !graph_pdl_suppressed()) ? 1 : 0;
  PKB_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
// fast-intrinsic forms (ex2.approx + rcp.approx, ~2 ulp): these sit in GEMM epilogues where an IEEE division would dominate
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Split an f32 into bf16 hi + bf16 lo (hi = rn(x), lo = rn(x - hi)): hi+lo carries ~16 mantissa bits.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// Store one GEMM-A-operand element.  Activations feeding a GEMM are bf16 [rows, lda]; in split ("precise") mode a
// second plane lo_off elements further holds the bf16 low parts (lo_off == 0: plain bf16).
__device__ __forceinline__ void store_act(__nv_bfloat16* A, size_t row, int lda, int col, float v, long long lo_off) {
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  A[row * (size_t)lda + col] = hi;
  if (lo_off) A[row * (size_t)lda + col + lo_off] = lo;
}
// 4 adjacent operand elements (col % 4 == 0): one 8-byte store per plane
__device__ __forceinline__ void store_act4(__nv_bfloat16* A, size_t row, int lda, int col, float4 v, long long lo_off) {
  __nv_bfloat16* dst = A + row * (size_t)lda + col;
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  if (lo_off) {
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    *reinterpret_cast<uint2*>(dst + lo_off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}
#endif

}  // namespace pkb
