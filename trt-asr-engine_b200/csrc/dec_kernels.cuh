// dec_kernels.cuh -- device-side state and kernels of the batched TDT greedy decode (see dec_kernels.cu).
#pragma once
#include "common.cuh"
#include "enc_kernels.cuh"
#include "gemm.h"

namespace pkb {

constexpr int kMaxStepsPerChunk = 32;   // streaming: >= valid_out_len * max_symbols (3*8 = 24 joint evaluations per chunk)
constexpr int kMaxStepsOffline = kMaxTq * (kMaxSymbols + 1);   // offline push: up to 32 encoder frames

struct DecodeDev {
  int B = 0;
  int max_symbols = kMaxSymbols;
  int max_steps = kMaxStepsPerChunk;   // capacity (records per entry) of `steps`
  int punct_suppress = 1;
  float blank_penalty = 0.0f;
  // per entry (this step)
  const int* slot = nullptr;      // [B]
  const int* row_off = nullptr;   // [B+1] packed-row prefix (encoder rows of entry e start at row_off[e])
  const int* t_enc = nullptr;     // [B] frames to decode = min(qlen, valid_out_len)
  int* t_cur = nullptr;           // [B]
  int* n_sym = nullptr;           // [B]
  int* active = nullptr;          // [B]
  int* emit_tok = nullptr;        // [B] token emitted in this iteration or -1
  int* pred_rowmap = nullptr;     // [B] slot if emitted else -1 (row map of the pred-projection epilogue)
  int* n_steps = nullptr;         // [B]
  int* steps = nullptr;           // [B][max_steps][3] = (time_idx, token, duration)
  int* n_active = nullptr;        // scalar: entries still active after the last select
  int* m_pred = nullptr;          // scalar: B if any entry emitted in this iteration else 0 (device-side GEMM M)
  int* m_joint = nullptr;         // scalar: B while any entry is still active at the start of the iteration else 0 (joint GEMM M):
                                  // iterations enqueued speculatively after the batch has finished cost nothing
  // tensors
  const float* enc_proj = nullptr;   // [M,640]  joint.enc(enc) + bias
  float* pred_proj = nullptr;        // [slots,640] joint.pred(g) + bias (cached per stream, refreshed on emission)
  const float* logits = nullptr;     // [B,8198]  (SIMT joint path: full logits)
  const float* part_val = nullptr;   // tensor-core joint path: [B][kArgmaxParts] slab maxima / first argmaxes of the token head
  const int* part_idx = nullptr;     //   (GEMM epilogue EPI_ARGMAX; NaN guard and blank penalty already applied)
  const float* dur_logits = nullptr; // [B][kNDur]
  int fused_argmax = 0;              // 1: select from the partials, 0: scan d.logits
  unsigned long long loop_handle = 0;   // != 0: the iteration is the body of a CUDA-graph WHILE node; its last-but-one kernel tells the
                                        // node whether any entry is still active (cudaGraphSetConditional)
  const float* gates = nullptr;      // [B,2560]
  const __nv_bfloat16* embed = nullptr;  // [8193,640]
  const unsigned* punct_bits = nullptr;  // [ceil(8193/32)]
  // per slot state
  float* pred_h = nullptr;        // [slots][2][640]
  float* pred_c = nullptr;
  float* pred_g = nullptr;        // [slots][640]
  int* n_emitted = nullptr;       // [slots] tokens emitted so far in the utterance
  int* y_id = nullptr;            // [slots] last emitted / primed token
  // GEMM operands
  ActOut act_hidden{};            // [B,640]
  ActOut act_pred{};              // [B,1280]
  ActOut act_g{};                 // [B,640]
};

#ifdef __CUDACC__
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
#endif

// Persistent decode loop (dec_kernels.cu): one cooperative kernel runs every pass of a whole-utterance decode of <= kPersistMaxB entries.
// Its pass costs ~10 + 1.45 B us (CUDA cores, every CTA meets every entry); the WHILE-graph body with tensor-core GEMMs ~40 us flat:
// measured 1.42 vs 3.24 s at 4 x 1 h, 5.09 vs 3.98 s at 32 x 1 h -- hence the limit.
constexpr int kPersistMaxB = 16;
struct DecPersistArgs {
  DecodeDev d;
  const __nv_bfloat16* w_out; const float* b_out;     // joint output layer [8198,640], [8198]
  const __nv_bfloat16* w_l0; const float* b_l0;       // LSTM layer 0 [2560,1280] = [weight_ih | weight_hh], [2560]
  const __nv_bfloat16* w_l1; const float* b_l1;
  const __nv_bfloat16* w_jp; const float* b_jp;       // joint.pred [640,640], [640]
  unsigned* bar;            // grid-barrier counter (zeroed before the launch)
  int* err;                 // set when a barrier times out (zeroed before the launch)
  int* flags;               // [2][2] per-pass (any emission, still-active count), zeroed before the launch
  float* part_val; int* part_idx;   // [B][grid] per-CTA (max, first argmax) of the token head
  float* dur;               // [B][5]
  float* xin; float* x1;    // [B][1280] LSTM layer inputs
  float* gvec;              // [B][640]
  int max_passes;
  int* passes_out;
};
size_t decode_persistent_smem();
bool launch_decode_persistent(const DecPersistArgs& a, int grid, cudaStream_t st);

void launch_decode_begin(const DecodeDev& d, cudaStream_t st);
void launch_decode_iter_reset(const DecodeDev& d, cudaStream_t st);
void launch_joint_hidden(const DecodeDev& d, cudaStream_t st);
void launch_tdt_select(const DecodeDev& d, cudaStream_t st);
void launch_pred_input(const DecodeDev& d, cudaStream_t st);
void launch_lstm_cell(const DecodeDev& d, int layer, cudaStream_t st);
void launch_force_token(const DecodeDev& d, const int* d_toks, cudaStream_t st);

}  // namespace pkb
