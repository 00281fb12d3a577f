// gemm.h -- dense projections C[m,n] = sum_k A[m,k] * W[n,k] with fused epilogues.
//
// A: bf16 activations [M, lda] row-major.  a_lo_off == 0: plain bf16.  a_lo_off != 0 ("split"/precise mode): a second
//    plane a_lo_off elements after A holds the bf16 LOW parts of the f32 activations (hi = rn(x), lo = rn(x-hi)),
//    so A*W is evaluated to ~fp32 accuracy with two bf16 tensor-core passes over the same weight tile.
// W: bf16 weights [N, K] row-major (PyTorch Linear layout), both operands K-major.
//
// Two backends share the epilogues:
//   * gemm_tc   (gemm_tc.cu)   tcgen05.mma + TMA + TMEM, 128x128x64 tiles      -- the throughput path
//   * gemm_simt (gemm_simt.cu) weight-streaming warp-per-column kernel on CUDA cores -- M <= 16 latency path
//                              and the always-correct fallback the tensor-core path is validated against.
#pragma once
#include "common.cuh"

namespace pkb {

enum EpiMode : int {
  EPI_BIAS_F32 = 0,        // out_f32[m,n] = acc + bias[n]
  EPI_BIAS_RELU_F32 = 1,   // out_f32[m,n] = relu(acc + bias[n])
  EPI_BIAS_RELU_ACT = 2,   // out_act[m,n] = relu(acc + bias[n])           (bf16 hi[/lo])
  EPI_BIAS_ROWMAP_F32 = 3, // out_f32[row_map[m],n] = acc + bias[n]        (row_map[m] < 0: dropped)
  EPI_SILU_ACT = 4,        // out_act[m,n] = silu(acc)
  EPI_RESADD_F32 = 5,      // out_f32[m,n] += scale * acc
  EPI_QKV = 6,             // n<1024: q f32; [1024,2048): K^T ring; [2048,3072): V ring
  EPI_GLU_F32 = 7,         // interleaved weights: out_f32[m,n/2] = acc[n] * sigmoid(acc[n+1]), n even
  EPI_F32 = 8,             // out_f32[m,n] = acc
  EPI_PARTIAL_F32 = 10,
  EPI_ACT = 11,            // out_act[m,n] = acc   (bf16 hi[/lo], no bias / activation)
  EPI_ARGMAX = 9,          // joint output layer with the greedy selection fused (tensor-core backend only): per row and per
                           // 128-column slab the running (max, first argmax) of acc + bias over the token head [0, kVocab)
                           // goes to part_val / part_idx [m][2 * tiles_n]; the kNDur duration logits go to dur_out [m][kNDur].
                           // NaN -> -100 and the blank penalty are applied here; the logits are never written.
};
// EPI_PARTIAL_F32 (= 10, tensor-core backend only): split-K.  The k-range is cut into epi.splits parts, split s writes its raw
// partial sums to out_f32[(s * part_rows + m) * ldo + n]; the consumer (the LayerNorm that follows every residual GEMM) adds
// them to the residual stream in a fixed order, so the result does not depend on scheduling.
constexpr int kArgmaxParts = 2 * ((kJointOut + 255) / 256);   // slabs per row (256-wide tiles, two 128-column halves each)

struct EpiParams {
  int mode = EPI_F32;
  float* out_f32 = nullptr;
  int ldo = 0;
  __nv_bfloat16* out_act = nullptr;
  int lda_out = 0;          // row stride of out_act
  long long lo_off_out = 0; // element offset of out_act's lo plane (0: plain bf16)
  int n_off = 0;            // added to n before the mode logic (used when W points into a row range of a fused matrix)
  const float* bias = nullptr;
  const int* row_map = nullptr;
  float scale = 1.0f;
  // EPI_QKV
  const int* row_entry = nullptr;   // [M] batch entry of each packed row
  const int* row_pos = nullptr;     // [M] position of the row inside its chunk
  const int* entry_slot = nullptr;  // [B]
  const int* entry_head = nullptr;  // [B] physical ring index of logical cache position 0
  void* kring = nullptr;            // this layer's K^T ring  [slot][head][d][kRingCap]
  void* vring = nullptr;            // this layer's V ring    [slot][kRingCap][1024]
  int kv_f32 = 0;                   // ring element type: 1 = f32 (precise mode), 0 = bf16
  __nv_bfloat16* q_bf16 = nullptr;  // non-null: q leaves as two bf16 planes (q + pos_bias_u, q + pos_bias_v) instead of f32
  long long q_plane = 0;            // elements between the planes
  const float* bias_u = nullptr;    // [1024] = pos_bias_u[head][128] flattened
  const float* bias_v = nullptr;
  int splits = 1;                   // EPI_PARTIAL_F32: number of k-splits
  int part_rows = 0;                //                  rows per split in the workspace
  int part_bf16 = 0;                //                  1: partial sums stored as bf16 (bf16 mode; halves the workspace traffic)
  int pair_split = 0;               //                  1: the split count was chosen for the CTA-pair kernel (256 x 256 tiles); 2: ... and the
                                    //                  pair kernel is to be used whatever its shape heuristic says (small batches, long K)
  int part_wide = 0;                //                  1: 128 x 256 tiles (single-CTA kernel) instead of 128 x 128: fewer operand bytes per MAC
  float* part_val = nullptr;        // EPI_ARGMAX: [M][kArgmaxParts]
  int* part_idx = nullptr;
  float* dur_out = nullptr;         // [M][kNDur]
  float blank_penalty = 0.0f;
  int direct_bf16 = 0;              // set by gemm_tc(): bf16 output rows are written straight from the TMEM-load registers (see gemm_tc.cu)
  int k_natural = 0;                // 1: K and V rings head-major [slot][head][kRingCap][128] (tensor-core attention);
                                    // 0: K^T ring [slot][head][128][kRingCap] + V ring [slot][kRingCap][1024] (precise mode)
};

struct GemmArgs {
  const __nv_bfloat16* A = nullptr;
  int lda = 0;
  long long a_lo_off = 0;
  const __nv_bfloat16* W = nullptr;
  int M = 0, N = 0, K = 0;
  // batched problems sharing one launch (tensor-core backend): batch b multiplies A[:, b*a_col_stride : +K] by
  // W[b*w_row_stride : +N, :] and writes at column offset b*out_col_stride of the output (e.g. one batch per attention head)
  int batch = 1;
  int a_col_stride = 0, w_row_stride = 0, out_col_stride = 0;
  const int* M_dev = nullptr;   // optional device-side row count: effective M = min(*M_dev, M) (decode: rows known only on device)
  const struct TensorMap* map_w32 = nullptr;   // host pointer, optional: W's tensor map with a 32-row box (CTA-pair kernel: column slices
                                               // of the tiles of a partly filled last round)
  EpiParams epi;
};

#ifdef __CUDACC__
// Epilogue on a pair of adjacent columns (n even).  Shared by both backends.
__device__ __forceinline__ void epilogue_pair(const EpiParams& p, int m, int n, int N, float v0, float v1) {
  const bool has1 = (n + 1) < N;
  n += p.n_off;
  switch (p.mode) {
    case EPI_F32:
      p.out_f32[(size_t)m * p.ldo + n] = v0;
      if (has1) p.out_f32[(size_t)m * p.ldo + n + 1] = v1;
      break;
    case EPI_BIAS_F32:
      p.out_f32[(size_t)m * p.ldo + n] = v0 + p.bias[n];
      if (has1) p.out_f32[(size_t)m * p.ldo + n + 1] = v1 + p.bias[n + 1];
      break;
    case EPI_BIAS_RELU_F32:
      p.out_f32[(size_t)m * p.ldo + n] = fmaxf(v0 + p.bias[n], 0.0f);
      if (has1) p.out_f32[(size_t)m * p.ldo + n + 1] = fmaxf(v1 + p.bias[n + 1], 0.0f);
      break;
    case EPI_BIAS_RELU_ACT:
      store_act(p.out_act, m, p.lda_out, n, fmaxf(v0 + p.bias[n], 0.0f), p.lo_off_out);
      if (has1) store_act(p.out_act, m, p.lda_out, n + 1, fmaxf(v1 + p.bias[n + 1], 0.0f), p.lo_off_out);
      break;
    case EPI_BIAS_ROWMAP_F32: {
      const int r = p.row_map[m];
      if (r >= 0) {
        p.out_f32[(size_t)r * p.ldo + n] = v0 + p.bias[n];
        if (has1) p.out_f32[(size_t)r * p.ldo + n + 1] = v1 + p.bias[n + 1];
      }
      break;
    }
    case EPI_ACT:
      store_act(p.out_act, m, p.lda_out, n, v0, p.lo_off_out);
      if (has1) store_act(p.out_act, m, p.lda_out, n + 1, v1, p.lo_off_out);
      break;
    case EPI_SILU_ACT:
      store_act(p.out_act, m, p.lda_out, n, silu(v0), p.lo_off_out);
      if (has1) store_act(p.out_act, m, p.lda_out, n + 1, silu(v1), p.lo_off_out);
      break;
    case EPI_RESADD_F32: {
      float* o = p.out_f32 + (size_t)m * p.ldo + n;
      o[0] += p.scale * v0;
      if (has1) o[1] += p.scale * v1;
      break;
    }
    case EPI_GLU_F32:
      if (p.out_act) p.out_act[(size_t)m * p.lda_out + (n >> 1)] = __float2bfloat16_rn(v0 * sigmoidf_(v1));
      else p.out_f32[(size_t)m * p.ldo + (n >> 1)] = v0 * sigmoidf_(v1);
      break;
    case EPI_QKV: {
      if (n < kDModel && p.q_bf16) {
        __nv_bfloat16* q = p.q_bf16 + (size_t)m * kDModel + n;
        q[0] = __float2bfloat16_rn(v0 + p.bias_u[n]); q[1] = __float2bfloat16_rn(v1 + p.bias_u[n + 1]);
        q[p.q_plane] = __float2bfloat16_rn(v0 + p.bias_v[n]); q[p.q_plane + 1] = __float2bfloat16_rn(v1 + p.bias_v[n + 1]);
      } else if (n < kDModel) {
        p.out_f32[(size_t)m * p.ldo + n] = v0;
        p.out_f32[(size_t)m * p.ldo + n + 1] = v1;
      } else {
        const int e = p.row_entry[m];
        const int slot = p.entry_slot[e];
        const int phys = (p.entry_head[e] + kCacheS + p.row_pos[m]) % kRingCap;
        if (p.k_natural) {      // bf16 mode: both rings head-major [slot][head][kRingCap][128]
          const int c = (n - kDModel) & (kDModel - 1), h = c >> 7, d = c & 127;
          const size_t i0 = (((size_t)slot * kHeads + h) * kRingCap + phys) * kDHead + d;
          void* ring = n < 2 * kDModel ? p.kring : p.vring;
          if (p.kv_f32) { ((float*)ring)[i0] = v0; ((float*)ring)[i0 + 1] = v1; }
          else { ((__nv_bfloat16*)ring)[i0] = __float2bfloat16_rn(v0); ((__nv_bfloat16*)ring)[i0 + 1] = __float2bfloat16_rn(v1); }
        } else if (n < 2 * kDModel) {
          const int c = n - kDModel, h = c >> 7, d = c & 127;
          const size_t i0 = (((size_t)slot * kHeads + h) * kDHead + d) * kRingCap + phys;
          if (p.kv_f32) { ((float*)p.kring)[i0] = v0; ((float*)p.kring)[i0 + kRingCap] = v1; }
          else { ((__nv_bfloat16*)p.kring)[i0] = __float2bfloat16_rn(v0); ((__nv_bfloat16*)p.kring)[i0 + kRingCap] = __float2bfloat16_rn(v1); }
        } else {
          const size_t i0 = ((size_t)slot * kRingCap + phys) * kDModel + (n - 2 * kDModel);
          if (p.kv_f32) { ((float*)p.vring)[i0] = v0; ((float*)p.vring)[i0 + 1] = v1; }
          else { ((__nv_bfloat16*)p.vring)[i0] = __float2bfloat16_rn(v0); ((__nv_bfloat16*)p.vring)[i0 + 1] = __float2bfloat16_rn(v1); }
        }
      }
      break;
    }
    default: break;
  }
}
#endif

// CUDA-core backend (any M; efficient for M <= 16).
void gemm_simt(const GemmArgs& g, cudaStream_t st);
// ... with the LayerNorm that produces A folded in (A = LN(x) computed per CTA; g.A is not read).  K == 1024, bf16 mode, M <= 16.
struct LnFuse;
bool gemm_simt_ln_supported(const GemmArgs& g);
void gemm_simt_ln(const GemmArgs& g, const LnFuse& f, cudaStream_t st);

// tcgen05 backend.  `map_a` / `map_w` are CUtensorMap objects (128-byte opaque, 64-byte aligned) created by
// make_tensor_map_2d for A [rows, lda] (hi plane then lo plane: lo rows start at a_lo_off / lda) and W [N, K],
// both with a {64, 128} box and 128-byte swizzle.
struct TensorMap { alignas(64) unsigned char bytes[128]; };
void make_tensor_map_2d(TensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                        uint32_t box_rows);
void gemm_tc(const GemmArgs& g, const TensorMap& map_a, const TensorMap& map_w, cudaStream_t st);
bool gemm_tc_supported(const GemmArgs& g);
int gemm_tc_pair_splits(int M, int N, int K);   // 0: pair kernel not used for this shape; else recommended k-splits (1 or 2)
void gemm_tc_set_bn(int bn);   // validation hook: 0 = heuristic tile width, 128 / 256 = forced

}  // namespace pkb
