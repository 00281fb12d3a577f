// dec_kernels.cu -- batched TDT greedy decode (piece 3): the non-GEMM kernels around the predictor / joint GEMMs.
//
// Control flow restated from /root/reference/cpp/src/parakeet_trt.cpp:2914-3676 (== tools/verify_nemo/tdt_trace.py:277-353):
// per encoder frame up to 8 symbols; token = first-max argmax over [0,8193); duration = argmax over [8193,8198);
// blank with duration 0 advances by 1; a non-blank token runs the predictor; advance 0 stays on the frame; after 8
// symbols without advance the frame is forced forward; an advance past the chunk end is dropped.  The reference runs
// this loop on the host with two device syncs per symbol; here every stream of the batch moves one symbol per
// "iteration" entirely on the device, the host only polls a single counter to learn when all streams are done.
#include "dec_kernels.cuh"

namespace pkb {

// hidden[e] = relu(E[row(e, t_e)] + P[slot_e])  -> GEMM operand rows [B,640]  (E, P already include their biases)
__global__ void __launch_bounds__(128)
joint_hidden_kernel(DecodeDev d) {
  pdl_enter();
  const int e = blockIdx.x;
  const bool act = d.active[e] != 0;
  const float* E = d.enc_proj + (size_t)(d.row_off[e] + (act ? d.t_cur[e] : 0)) * kJointH;
  const float* P = d.pred_proj + (size_t)d.slot[e] * kJointH;
  for (int c = threadIdx.x; c < kJointH; c += 128) {
    const float v = act ? fmaxf(E[c] + P[c], 0.0f) : 0.0f;
    store_act(d.act_hidden.ptr, e, d.act_hidden.lda, c, v, d.act_hidden.lo_off);
  }
}

// One CTA per entry: fused argmax over both heads + the TDT advance rules + state update.
__global__ void __launch_bounds__(256)
tdt_select_kernel(DecodeDev d) {
  pdl_enter();
  const int e = blockIdx.x, tid = threadIdx.x;
  if (!d.active[e]) {
    if (tid == 0) { d.emit_tok[e] = -1; d.pred_rowmap[e] = -1; }
    return;
  }
  const float* lg = d.logits + (size_t)e * kJointOut;
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  if (d.fused_argmax) {
    // the joint GEMM's epilogue already reduced every 128-column slab: finish over the kArgmaxParts slab winners
    if (tid < kArgmaxParts) { best = d.part_val[(size_t)e * kArgmaxParts + tid]; bidx = d.part_idx[(size_t)e * kArgmaxParts + tid]; }
  } else {
    for (int i = tid; i < kVocab; i += 256) {
      float v = lg[i];
      if (v != v) v = -100.0f;                                  // NaN logits -> -100 (parakeet_trt.cpp:2971)
      if (i == kBlank) v -= d.blank_penalty;                    // PARAKEET_BLANK_PENALTY (:3175-3178), default 0
      if (v > best) { best = v; bidx = i; }                     // ascending i per thread: first max wins
    }
  }
  // block reduce (value desc, index asc) == "first maximum wins" of the reference's strict '>' scan (:3200-3206)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  if ((tid & 31) == 0) { s_v[tid >> 5] = best; s_i[tid >> 5] = bidx; }
  __syncthreads();
  if (tid != 0) return;
  for (int w = 1; w < 8; ++w)
    if (s_v[w] > best || (s_v[w] == best && s_i[w] < bidx)) { best = s_v[w]; bidx = s_i[w]; }
  int tok = bidx;
  const int slot = d.slot[e];
  // leading punctuation-only suppression while nothing has been emitted in this utterance (:3256-3262)
  if (d.punct_suppress && d.n_emitted[slot] == 0 && tok < kBlank && ((d.punct_bits[tok >> 5] >> (tok & 31)) & 1u)) tok = kBlank;
  const float* dl = d.fused_argmax ? d.dur_logits + (size_t)e * kNDur : lg + kVocab;
  int dbest = 0;
  float dv = dl[0];
  if (dv != dv) dv = -100.0f;
  for (int i = 1; i < kNDur; ++i) {
    float v = dl[i];
    if (v != v) v = -100.0f;
    if (v > dv) { dv = v; dbest = i; }
  }
  const int dur = dbest;                                      // duration_values = [0,1,2,3,4]
  const int adv = (tok == kBlank && dur == 0) ? 1 : dur;      // blank must advance (:3393-3403)
  int t = d.t_cur[e], ns = d.n_sym[e];
  const int k = d.n_steps[e];
  if (k < d.max_steps) {
    int* st = d.steps + ((size_t)e * d.max_steps + k) * 3;
    st[0] = t; st[1] = tok; st[2] = dur;
    d.n_steps[e] = k + 1;
  }
  if (tok != kBlank) {
    d.n_emitted[slot] += 1;
    d.y_id[slot] = tok;
    d.emit_tok[e] = tok;
    d.pred_rowmap[e] = slot;
    atomicMax(d.m_pred, d.B);
  } else {
    d.emit_tok[e] = -1;
    d.pred_rowmap[e] = -1;
  }
  ns += 1;
  if (adv == 0) {
    if (ns >= d.max_symbols) { t += 1; ns = 0; }             // forced advance (:3665-3676)
  } else {
    t += adv;
    ns = 0;
  }
  d.t_cur[e] = t;
  d.n_sym[e] = ns;
  if (t >= d.t_enc[e]) d.active[e] = 0;
  else atomicAdd(d.n_active, 1);
}

// act rows [emb(tok) ; h_layer0]  (K = 1280) for the layer-0 gate GEMM
__global__ void __launch_bounds__(128)
pred_input_kernel(DecodeDev d) {
  pdl_enter();
  const int e = blockIdx.x;
  if (*d.m_pred == 0) return;
  const int tok = d.emit_tok[e];
  const float* h0 = d.pred_h + (size_t)d.slot[e] * kPredL * kPredH;
  for (int c = threadIdx.x; c < 2 * kPredH; c += 128) {
    float v = 0.0f;
    if (tok >= 0) v = c < kPredH ? __bfloat162float(d.embed[(size_t)tok * kPredH + c]) : h0[c - kPredH];
    store_act(d.act_pred.ptr, e, d.act_pred.lda, c, v, d.act_pred.lo_off);
  }
}

// LSTM cell (gate order i,f,g,o) for one layer; writes the next GEMM operand.
//  layer 0: act rows <- [h0_new ; h1_old]       layer 1: act_g rows <- h1_new (K = 640), g state updated
__global__ void __launch_bounds__(128)
lstm_cell_kernel(DecodeDev d, int layer) {
  pdl_enter();
  const int e = blockIdx.x;
  // graph mode: this launch follows tdt_select_kernel in stream order, so the count of still-active entries is final here
  if (d.loop_handle != 0 && layer == 1 && e == 0 && threadIdx.x == 0)
    cudaGraphSetConditional((cudaGraphConditionalHandle)d.loop_handle, *d.n_active > 0 ? 1u : 0u);
  if (*d.m_pred == 0) return;
  const int tok = d.emit_tok[e];
  const int slot = d.slot[e];
  float* h = d.pred_h + ((size_t)slot * kPredL + layer) * kPredH;
  float* c = d.pred_c + ((size_t)slot * kPredL + layer) * kPredH;
  const float* gates = d.gates + (size_t)e * 4 * kPredH;
  for (int j = threadIdx.x; j < kPredH; j += 128) {
    float hn = 0.0f;
    if (tok >= 0) {
      const float ig = sigmoidf_acc(gates[j]);
      const float fg = sigmoidf_acc(gates[kPredH + j]);
      const float gg = tanhf(gates[2 * kPredH + j]);
      const float og = sigmoidf_acc(gates[3 * kPredH + j]);
      const float cn = fg * c[j] + ig * gg;
      hn = og * tanhf(cn);
      c[j] = cn;
      h[j] = hn;
    }
    if (layer == 0) {
      store_act(d.act_pred.ptr, e, d.act_pred.lda, j, hn, d.act_pred.lo_off);
      const float h1 = tok >= 0 ? d.pred_h[((size_t)slot * kPredL + 1) * kPredH + j] : 0.0f;
      store_act(d.act_pred.ptr, e, d.act_pred.lda, kPredH + j, h1, d.act_pred.lo_off);
    } else {
      if (tok >= 0) d.pred_g[(size_t)slot * kPredH + j] = hn;
      store_act(d.act_g.ptr, e, d.act_g.lda, j, hn, d.act_g.lo_off);
    }
  }
}

__global__ void decode_begin_kernel(DecodeDev d) {
  pdl_enter();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) { *d.n_active = 1; *d.m_pred = 0; *d.m_joint = 0; }      // n_active > 0: the first iteration runs
  if (e >= d.B) return;
  d.t_cur[e] = 0;
  d.n_sym[e] = 0;
  d.n_steps[e] = 0;
  d.active[e] = d.t_enc[e] > 0 ? 1 : 0;
  d.emit_tok[e] = -1;
  d.pred_rowmap[e] = -1;
}

__global__ void decode_iter_reset_kernel(DecodeDev d) {
  pdl_enter();
  if (threadIdx.x == 0) { *d.m_joint = *d.n_active > 0 ? d.B : 0; *d.n_active = 0; *d.m_pred = 0; }
}

// priming / forced token (reset_utterance): mark entries to run the predictor on a given token
__global__ void force_token_kernel(DecodeDev d, const int* toks) {
  pdl_enter();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.B) return;
  d.emit_tok[e] = toks[e];
  d.pred_rowmap[e] = toks[e] >= 0 ? d.slot[e] : -1;
  if (toks[e] >= 0) atomicMax(d.m_pred, d.B);
}

void launch_decode_begin(const DecodeDev& d, cudaStream_t st) {
  if (d.B <= 0) return;
  launch_k(decode_begin_kernel, dim3((d.B + 127) / 128), dim3(128), 0, st, d);
  PKB_CUDA(cudaGetLastError());
}
void launch_decode_iter_reset(const DecodeDev& d, cudaStream_t st) {
  launch_k(decode_iter_reset_kernel, dim3(1), dim3(32), 0, st, d);
  PKB_CUDA(cudaGetLastError());
}
void launch_joint_hidden(const DecodeDev& d, cudaStream_t st) {
  if (d.B <= 0) return;
  launch_k(joint_hidden_kernel, dim3(d.B), dim3(128), 0, st, d);
  PKB_CUDA(cudaGetLastError());
}
void launch_tdt_select(const DecodeDev& d, cudaStream_t st) {
  if (d.B <= 0) return;
  launch_k(tdt_select_kernel, dim3(d.B), dim3(256), 0, st, d);
  PKB_CUDA(cudaGetLastError());
}
void launch_pred_input(const DecodeDev& d, cudaStream_t st) {
  if (d.B <= 0) return;
  launch_k(pred_input_kernel, dim3(d.B), dim3(128), 0, st, d);
  PKB_CUDA(cudaGetLastError());
}
void launch_lstm_cell(const DecodeDev& d, int layer, cudaStream_t st) {
  if (d.B <= 0) return;
  launch_k(lstm_cell_kernel, dim3(d.B), dim3(128), 0, st, d, layer);
  PKB_CUDA(cudaGetLastError());
}
void launch_force_token(const DecodeDev& d, const int* d_toks, cudaStream_t st) {
  if (d.B <= 0) return;
  launch_k(force_token_kernel, dim3((d.B + 127) / 128), dim3(128), 0, st, d, d_toks);
  PKB_CUDA(cudaGetLastError());
}

}  // namespace pkb
