// offline_long.cuh -- kernels of the whole-utterance ("long-form") offline encoder: full self-attention over ALL frames of an
// utterance, no caches.  This is what the reference's non-streaming `encoder` graph computes when it is run with a dynamic
// time axis (contracts/parakeet-tdt-0.6b-v3.contract.json:67-96; BASELINE configs 1 and 5: 10 s and 1 h clips), i.e. NeMo's
// ConformerEncoder.forward with att_context [-1,-1]: RelPositionalEncoding over 2T-1 relative positions,
// RelPositionMultiHeadAttention with rel_shift, symmetric (4,4) padding in the depthwise conv.
// The dense projections and LayerNorms are the streaming path's kernels; only the three pieces that look across time differ.
#pragma once
#include "enc_kernels.cuh"

namespace pkb {

// Sinusoidal relative-position table as a GEMM-A operand: row r <-> relative position (r - (Tm-1)), r in [0, 2Tm-1);
// pe[r][2i] = sin(pos * div_i), pe[r][2i+1] = cos(pos * div_i), div_i = exp(-(ln 1e4) * 2i / 1024).
void launch_lf_posemb(ActOut a, int Tm, cudaStream_t st);

struct LfAttnArgs {
  const __nv_bfloat16* qkv_bf16 = nullptr;   // bf16 mode:    [M][3072] = q | k | v, head h at columns h*128 inside each third
  const float* qkv_f32 = nullptr;            // precise mode: same layout, f32
  const __nv_bfloat16* ppos_bf16 = nullptr;  // projected position table linear_pos(pe): [2Tm-1][1024], row = rel + (Tm-1)
  const float* ppos_f32 = nullptr;
  int Tm = 0;                                // the table's centre: rows cover relative positions -(Tm-1) .. Tm-1
  int max_T = 0;                             // longest utterance of the batch (grid size)
  const float* bias_u = nullptr;             // [1024]
  const float* bias_v = nullptr;
  ActOut ctx{};                              // [M][1024] operand of linear_out
  // bf16 mode, TMA-staged kernel: host pointers to 128-byte CUtensorMap objects (gemm.h make_tensor_map_2d: 64-column boxes,
  // 128-byte swizzle) over qkv [M rows][3072] with 64-row boxes and over the table [2Tm-1 rows][1024] with 128-row boxes;
  // rows outside a map read as zero.  NULL: the LDG/STS-staged kernel runs instead.
  const void* map_qkv = nullptr;
  const void* map_pos = nullptr;
};
// Every entry e of the batch is an utterance of b.Tq[e] encoder rows starting at packed row b.row_off[e]; each query row
// attends to all rows of its utterance:  softmax_j(((q_i+u).k_j + (q_i+v).p_{i-j}) / sqrt(128)) v_j.
// bf16 mode: flash-style mma.sync kernel (64-query x 64-key tiles, online softmax, position term on the tensor cores from a
// 127-row window of the table).  precise mode: f32 CUDA-core kernel (parity-grade; scores of one query row live in smem).
void launch_lf_attention(const BatchDev& b, const LfAttnArgs& a, cudaStream_t st);

// tcgen05 version (lf_attn_tc.cu, bf16 mode).  lf_prep_kernel derives its operands from the layer's q | k | v rows once per layer:
// the two biased query planes and V transposed ([1024][ldv], utterance e at the 64-aligned column offset sum_{x<e} roundup64(T_x)).
struct LfTcArgs {
  const __nv_bfloat16* qkv = nullptr;      // [M][3072]
  const float* bias_u = nullptr;           // [1024]
  const float* bias_v = nullptr;
  __nv_bfloat16* q_planes = nullptr;       // [2][q_plane_rows][1024]: q + pos_bias_u | q + pos_bias_v
  long long q_plane = 0;                   // elements between the planes
  int q_plane_rows = 0;                    // rows between the planes
  __nv_bfloat16* vt = nullptr;             // [1024][ldv]
  long long ldv = 0;
  int Tm = 0;                              // centre of the projected position table (row = relative position + Tm - 1)
  __nv_bfloat16* ctx = nullptr;            // [M][ldc] operand of linear_out
  int ldc = 0;
  // host pointers to CUtensorMap objects (64-column boxes, 128-byte swizzle): query planes (128-row boxes), q|k|v rows (64-row boxes),
  // projected table (64-row boxes), V^T (128-row boxes)
  const void* map_q = nullptr;
  const void* map_k = nullptr;
  const void* map_pos = nullptr;
  const void* map_vt = nullptr;
  int wait_hint_ns = 0;                    // > 0: suspend-time hint of the single-thread roles' mbarrier polls (set by the launcher)
};
void launch_lf_prep(const BatchDev& b, const LfTcArgs& a, cudaStream_t st);
void launch_lf_attention_tc(const BatchDev& b, const LfTcArgs& a, int max_T, cudaStream_t st);

struct LfDwConvArgs {
  const float* c = nullptr;               // post-GLU activations f32 [M][1024] (precise mode), or
  const __nv_bfloat16* c_bf16 = nullptr;  // bf16 [M][1024] (bf16 mode)
  const float* w = nullptr;               // [1024][9], BatchNorm scale folded in
  const float* bias = nullptr;            // [1024] folded BatchNorm offset
  ActOut out{};                           // [M][1024]
  int M = 0;
};
// depthwise k=9 over the utterance with zero padding (4,4), folded BatchNorm, SiLU
void launch_lf_dwconv(const BatchDev& b, const LfDwConvArgs& a, cudaStream_t st);

}  // namespace pkb
