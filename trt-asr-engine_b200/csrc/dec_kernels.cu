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


// ================================================================================================ persistent decode loop
// Whole-utterance decode of a few utterances (<= kPersistMaxB): tens of thousands of passes, each a matrix-vector-sized problem.
// As a chain of launches (or as the body of a WHILE graph node) a pass costs ~10 kernel boundaries = 32 - 39 us whatever the batch
// (3.2 of the 5.7 s of BASELINE config 5's shape, 4 x 1 h).  Here ONE cooperative kernel runs the whole loop: a CTA per SM keeps ITS
// slice of every decoder weight in shared memory for the lifetime of the loop -- 56 rows of the joint output layer [8198,640], the
// four gate rows of 5 hidden units of each LSTM layer [2560,1280], 5 rows of joint.pred [640,640]: 180 KB per CTA, 24.5 MB over the
// chip -- and the passes are separated by grid-wide barriers (a monotonic counter in global memory):
//   phase 1  hidden = relu(E[t] + P) (every CTA, from L2), logits of the CTA's columns, per-entry (max, first argmax) -> global
//   phase 2  CTA e finishes entry e: argmax over the CTAs' partials, TDT advance rules (tdt_select_kernel's), trace record,
//            predictor input [emb(token); h0] for emitting entries
//   phase 3/4  LSTM layers (only when some entry emitted): gate rows of the CTA's units, cell update, state + next input
//   phase 5  joint.pred rows of the CTA -> P
// Activations stay f32 (the launch-chain path feeds bf16 hi + lo planes: 16 mantissa bits), weights bf16, f32 accumulation.
namespace {
constexpr int kPdThreads = 256, kPdWarps = kPdThreads / 32;
constexpr int kPdChunk = 8;                                     // entries processed together (register accumulators)
constexpr int kPdColsMax = 56, kPdUnitsMax = 5;                 // per-CTA slice capacity: a grid of >= 147 CTAs (8198 / 56, 640 / 5)
constexpr size_t kPdSmem = (size_t)kPdColsMax * kJointH * 2 + 2 * (size_t)kPdUnitsMax * 4 * 2 * kPredH * 2 + (size_t)kPdUnitsMax * kPredH * 2 +
                           (size_t)kPdChunk * 2 * kPredH * 4;      // 221 440 B of the 227 KB a CTA may own

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// grid barrier number `k` (1-based): every CTA adds 1 to a monotonic counter and waits until k * G arrivals have been seen
__device__ __forceinline__ bool pd_grid_sync(unsigned* bar, unsigned target, int* err) {
  __syncthreads();
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    int ok = 1;
    long long spins = 0;
    while (ld_acquire_u32(bar) < target) {
      if (++spins > (1ll << 27)) { ok = 0; atomicExch(err, 1); break; }      // watchdog: never hang the device
      if ((spins & 1023) == 0 && *reinterpret_cast<volatile int*>(err) != 0) { ok = 0; break; }
    }
    __threadfence();
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}
// Dot products of NR bf16 weight rows (shared memory; rows r0, r0 + rstep, ... below nrows) with the first NE of up to 8 f32 vectors
// (shared memory, pitch ldx), K % 128 == 0.  The activations of a k-step are loaded ONCE per lane and reused for every row (one row
// at a time re-read them per row: 8 LDS.128 per 32 FMAs, bound by the shared-memory pipe).  The 8 per-lane partial sums of a row
// are reduced "transposed": after the xor-16 / 8 / 4 exchanges a lane keeps the sum of ONE entry, e = bits 4..2 of its lane index
// (9 shuffles per row instead of 40); every entry still goes through the same xor-16, 8, 4, 2, 1 butterfly as warp_sum(), and
// addition commutes, so the totals are bit-identical to the plain reduction and do not depend on NE.
template <int K, int NE, int NR>
__device__ __forceinline__ void pd_dot_rows_ne(const __nv_bfloat16* __restrict__ w, int ldw, int r0, int rstep, int nrows, const float* __restrict__ x,
                                               int ldx, int lane, float (&tot)[NR]) {
  float acc[NR][kPdChunk];
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int e = 0; e < kPdChunk; ++e) acc[i][e] = 0.0f;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 128) {
    float4 v[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) v[e] = *reinterpret_cast<const float4*>(x + (size_t)e * ldx + k0 + lane * 4);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const int r = r0 + i * rstep;
      if (r < nrows) {
        const uint2 raw = *reinterpret_cast<const uint2*>(w + (size_t)r * ldw + k0 + lane * 4);
        const float w0 = __uint_as_float(raw.x << 16), w1 = __uint_as_float(raw.x & 0xffff0000u);
        const float w2 = __uint_as_float(raw.y << 16), w3 = __uint_as_float(raw.y & 0xffff0000u);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          acc[i][e] = fmaf(w0, v[e].x, acc[i][e]); acc[i][e] = fmaf(w1, v[e].y, acc[i][e]);
          acc[i][e] = fmaf(w2, v[e].z, acc[i][e]); acc[i][e] = fmaf(w3, v[e].w, acc[i][e]);
        }
      }
    }
  }
  const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    float k4[4], k2[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {      // lanes with bit 4 clear keep entries 0..3, the others 4..7
      const float give = b4 ? acc[i][q] : acc[i][q + 4], keep = b4 ? acc[i][q + 4] : acc[i][q];
      k4[q] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float give = b3 ? k4[q] : k4[q + 2], keep = b3 ? k4[q + 2] : k4[q];
      k2[q] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
    const float give = b2 ? k2[0] : k2[1], keep = b2 ? k2[1] : k2[0];
    float t = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    tot[i] = t;      // the dot product of row r0 + i * rstep with entry (lane >> 2) & 7
  }
}
template <int K, int NR>
__device__ __forceinline__ void pd_dot_rows(const __nv_bfloat16* __restrict__ w, int ldw, int r0, int rstep, int nrows, const float* __restrict__ x, int ldx,
                                            int lane, int ne, float (&tot)[NR]) {
  if (ne <= 1) pd_dot_rows_ne<K, 1, NR>(w, ldw, r0, rstep, nrows, x, ldx, lane, tot);
  else if (ne <= 2) pd_dot_rows_ne<K, 2, NR>(w, ldw, r0, rstep, nrows, x, ldx, lane, tot);
  else if (ne <= 4) pd_dot_rows_ne<K, 4, NR>(w, ldw, r0, rstep, nrows, x, ldx, lane, tot);
  else pd_dot_rows_ne<K, 8, NR>(w, ldw, r0, rstep, nrows, x, ldx, lane, tot);
}
constexpr int kPdColsPerWarp = (kPdColsMax + kPdWarps - 1) / kPdWarps;          // 7
constexpr int kPdGateRowsPerWarp = (kPdUnitsMax * 4 + kPdWarps - 1) / kPdWarps; // 3
}  // namespace

__global__ void __launch_bounds__(kPdThreads, 1)
decode_persistent_kernel(DecPersistArgs a) {
  extern __shared__ __align__(16) uint8_t pd_raw[];
  const DecodeDev& d = a.d;
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cols_per = (kJointOut + G - 1) / G, units_per = (kPredH + G - 1) / G;
  const int c0 = cta * cols_per, ncols = max(0, min(kJointOut, c0 + cols_per) - c0);
  const int u0 = cta * units_per, nunits = max(0, min(kPredH, u0 + units_per) - u0);
  __nv_bfloat16* s_wout = reinterpret_cast<__nv_bfloat16*>(pd_raw);                       // [ncols][640]
  __nv_bfloat16* s_wl0 = s_wout + (size_t)kPdColsMax * kJointH;                           // [nunits][4][1280]
  __nv_bfloat16* s_wl1 = s_wl0 + (size_t)kPdUnitsMax * 4 * 2 * kPredH;
  __nv_bfloat16* s_wjp = s_wl1 + (size_t)kPdUnitsMax * 4 * 2 * kPredH;                    // [nunits][640]
  float* s_x = reinterpret_cast<float*>(s_wjp + (size_t)kPdUnitsMax * kPredH);            // [8][1280] activations of the current chunk
  __shared__ float s_red_v[kPdWarps][kPdChunk];
  __shared__ int s_red_i[kPdWarps][kPdChunk];
  __shared__ float s_gates[kPdUnitsMax * 4][kPdChunk];
  __shared__ int s_flag[kPdChunk];

  // ---- weights of this CTA -> shared memory (once)
  for (int i = tid; i < ncols * (kJointH / 8); i += kPdThreads) {
    const int r = i / (kJointH / 8), c = i % (kJointH / 8);
    reinterpret_cast<uint4*>(s_wout + (size_t)r * kJointH)[c] = __ldg(reinterpret_cast<const uint4*>(a.w_out + (size_t)(c0 + r) * kJointH) + c);
  }
  for (int i = tid; i < nunits * 4 * (2 * kPredH / 8); i += kPdThreads) {
    const int r = i / (2 * kPredH / 8), c = i % (2 * kPredH / 8);      // r = unit * 4 + gate
    const size_t grow = (size_t)((r & 3) * kPredH + u0 + (r >> 2)) * (2 * kPredH);
    reinterpret_cast<uint4*>(s_wl0 + (size_t)r * 2 * kPredH)[c] = __ldg(reinterpret_cast<const uint4*>(a.w_l0 + grow) + c);
    reinterpret_cast<uint4*>(s_wl1 + (size_t)r * 2 * kPredH)[c] = __ldg(reinterpret_cast<const uint4*>(a.w_l1 + grow) + c);
  }
  for (int i = tid; i < nunits * (kPredH / 8); i += kPdThreads) {
    const int r = i / (kPredH / 8), c = i % (kPredH / 8);
    reinterpret_cast<uint4*>(s_wjp + (size_t)r * kPredH)[c] = __ldg(reinterpret_cast<const uint4*>(a.w_jp + (size_t)(u0 + r) * kPredH) + c);
  }
  __syncthreads();

  unsigned nbar = 0;
  int pass = 0;
  for (; pass < a.max_passes; ++pass) {
    int* flags = a.flags + 2 * (pass & 1);      // [0] any entry emitted in this pass, [1] entries still active after it
    // ================= phase 1: joint output layer, columns [c0, c0 + ncols)
    for (int e0 = 0; e0 < d.B; e0 += kPdChunk) {
      const int ne = min(kPdChunk, d.B - e0);
      if (tid < kPdChunk) s_flag[tid] = (tid < ne && __ldcg(d.active + e0 + tid) != 0) ? 1 : 0;
      __syncthreads();
      bool any = false;
#pragma unroll
      for (int e = 0; e < kPdChunk; ++e) any |= s_flag[e] != 0;
      if (any) {
        for (int i = tid; i < kPdChunk * (kJointH / 4); i += kPdThreads) {
          const int e = i / (kJointH / 4), c = i % (kJointH / 4);
          float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
          if (s_flag[e]) {
            const int ge = e0 + e;
            const float4 E = __ldcg(reinterpret_cast<const float4*>(d.enc_proj + (size_t)(d.row_off[ge] + __ldcg(d.t_cur + ge)) * kJointH) + c);
            const float4 P = __ldcg(reinterpret_cast<const float4*>(d.pred_proj + (size_t)d.slot[ge] * kJointH) + c);
            h = make_float4(fmaxf(E.x + P.x, 0.f), fmaxf(E.y + P.y, 0.f), fmaxf(E.z + P.z, 0.f), fmaxf(E.w + P.w, 0.f));
          }
          reinterpret_cast<float4*>(s_x + (size_t)e * kJointH)[c] = h;
        }
        __syncthreads();
        // this lane's entry after the transposed reduction, and its running first maximum over the CTA's columns (ascending)
        const int my_e = (lane >> 2) & 7;
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        float tot[kPdColsPerWarp];
        pd_dot_rows<kJointH, kPdColsPerWarp>(s_wout, kJointH, warp, kPdWarps, ncols, s_x, kJointH, lane, ne, tot);
#pragma unroll
        for (int i = 0; i < kPdColsPerWarp; ++i) {
          const int j = warp + i * kPdWarps;
          if (j < ncols) {
            const int col = c0 + j;
            float y = tot[i] + __ldg(a.b_out + col);
            if (y != y) y = -100.0f;                                      // NaN logits -> -100 (parakeet_trt.cpp:2971)
            if (col == kBlank) y -= d.blank_penalty;                      // PARAKEET_BLANK_PENALTY (:3175-3178)
            if (col < kVocab) { if (y > best) { best = y; bidx = col; } }
            else if ((lane & 3) == 0 && my_e < ne) a.dur[(size_t)(e0 + my_e) * kNDur + (col - kVocab)] = y;
          }
        }
        if ((lane & 3) == 0) { s_red_v[warp][my_e] = best; s_red_i[warp][my_e] = bidx; }
        __syncthreads();
        if (tid < ne) {
          float bv = s_red_v[0][tid];
          int bi = s_red_i[0][tid];
          for (int w = 1; w < kPdWarps; ++w)
            if (s_red_v[w][tid] > bv || (s_red_v[w][tid] == bv && s_red_i[w][tid] < bi)) { bv = s_red_v[w][tid]; bi = s_red_i[w][tid]; }
          a.part_val[(size_t)(e0 + tid) * G + cta] = bv;
          a.part_idx[(size_t)(e0 + tid) * G + cta] = bi;
        }
      }
      __syncthreads();
    }
    if (!pd_grid_sync(a.bar, ++nbar * G, a.err)) return;

    // ================= phase 2: CTA e finishes entry e (greedy selection + TDT rules), predictor input for emitting entries
    if (cta == 0 && tid == 0) { a.flags[2 * ((pass + 1) & 1)] = 0; a.flags[2 * ((pass + 1) & 1) + 1] = 0; }
    for (int e = cta; e < d.B; e += G) {
      __shared__ int s_tok;
      if (tid == 0) s_tok = -1;
      if (__ldcg(d.active + e) != 0) {
        // thread 0 requests the scalars of the advance rules now, so that their round trips overlap the reduction below
        int p_slot = 0, p_nem = 0, p_t = 0, p_ns = 0, p_k = 0, p_tenc = 0;
        float p_dl[kNDur] = {0.f, 0.f, 0.f, 0.f, 0.f};
        if (tid == 0) {
          p_slot = d.slot[e]; p_t = d.t_cur[e]; p_ns = d.n_sym[e]; p_k = d.n_steps[e]; p_tenc = d.t_enc[e];
#pragma unroll
          for (int i = 0; i < kNDur; ++i) p_dl[i] = __ldcg(a.dur + (size_t)e * kNDur + i);
          p_nem = d.n_emitted[p_slot];
        }
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        for (int i = tid; i < G; i += kPdThreads) {
          const float v = __ldcg(a.part_val + (size_t)e * G + i);
          const int ix = __ldcg(a.part_idx + (size_t)e * G + i);
          if (v > best || (v == best && ix < bidx)) { best = v; bidx = ix; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
        }
        if (lane == 0) { s_red_v[warp][0] = best; s_red_i[warp][0] = bidx; }
        __syncthreads();
        if (tid == 0) {
          for (int w = 1; w < kPdWarps; ++w)
            if (s_red_v[w][0] > best || (s_red_v[w][0] == best && s_red_i[w][0] < bidx)) { best = s_red_v[w][0]; bidx = s_red_i[w][0]; }
          int tok = bidx;
          const int slot = p_slot;
          if (d.punct_suppress && p_nem == 0 && tok < kBlank && ((d.punct_bits[tok >> 5] >> (tok & 31)) & 1u)) tok = kBlank;
          int dbest = 0;
          float dv = p_dl[0];
          if (dv != dv) dv = -100.0f;
#pragma unroll
          for (int i = 1; i < kNDur; ++i) {
            float v = p_dl[i];
            if (v != v) v = -100.0f;
            if (v > dv) { dv = v; dbest = i; }
          }
          const int dur = dbest;
          const int adv = (tok == kBlank && dur == 0) ? 1 : dur;
          int t = p_t, ns = p_ns;
          const int k = p_k;
          if (k < d.max_steps) {
            int* st = d.steps + ((size_t)e * d.max_steps + k) * 3;
            st[0] = t; st[1] = tok; st[2] = dur;
            d.n_steps[e] = k + 1;
          }
          if (tok != kBlank) {
            d.n_emitted[slot] = p_nem + 1;
            d.y_id[slot] = tok;
            d.emit_tok[e] = tok;
            d.pred_rowmap[e] = slot;
            atomicExch(flags, 1);
            s_tok = tok;
          } else {
            d.emit_tok[e] = -1;
            d.pred_rowmap[e] = -1;
          }
          ns += 1;
          if (adv == 0) {
            if (ns >= d.max_symbols) { t += 1; ns = 0; }
          } else {
            t += adv;
            ns = 0;
          }
          d.t_cur[e] = t;
          d.n_sym[e] = ns;
          if (t >= p_tenc) d.active[e] = 0;
          else atomicAdd(flags + 1, 1);
        }
        __syncthreads();
        const int tok = s_tok;
        if (tok >= 0) {      // layer-0 input [emb(tok) ; h0]
          const float* h0 = d.pred_h + (size_t)d.slot[e] * kPredL * kPredH;
          for (int c = tid; c < 2 * kPredH; c += kPdThreads)
            a.xin[(size_t)e * 2 * kPredH + c] = c < kPredH ? __bfloat162float(d.embed[(size_t)tok * kPredH + c]) : __ldcg(h0 + c - kPredH);
        }
      } else if (tid == 0) {
        d.emit_tok[e] = -1;
        d.pred_rowmap[e] = -1;
      }
      __syncthreads();
    }
    if (!pd_grid_sync(a.bar, ++nbar * G, a.err)) return;
    const int any_emit = __ldcg(flags), n_active = __ldcg(flags + 1);

    if (any_emit) {
      // ================= phases 3 and 4: LSTM layers, units [u0, u0 + nunits) of this CTA
      for (int layer = 0; layer < kPredL; ++layer) {
        const __nv_bfloat16* s_w = layer == 0 ? s_wl0 : s_wl1;
        const float* bias = layer == 0 ? a.b_l0 : a.b_l1;
        const float* xsrc = layer == 0 ? a.xin : a.x1;
        for (int e0 = 0; e0 < d.B && nunits > 0; e0 += kPdChunk) {
          const int ne = min(kPdChunk, d.B - e0);
          if (tid < kPdChunk) s_flag[tid] = (tid < ne && __ldcg(d.emit_tok + e0 + tid) >= 0) ? 1 : 0;
          __syncthreads();
          bool any = false;
#pragma unroll
          for (int e = 0; e < kPdChunk; ++e) any |= s_flag[e] != 0;
          if (any) {
            for (int i = tid; i < kPdChunk * (2 * kPredH / 4); i += kPdThreads) {
              const int e = i / (2 * kPredH / 4), c = i % (2 * kPredH / 4);
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (s_flag[e]) v = __ldcg(reinterpret_cast<const float4*>(xsrc + (size_t)(e0 + e) * 2 * kPredH) + c);
              reinterpret_cast<float4*>(s_x + (size_t)e * 2 * kPredH)[c] = v;
            }
            __syncthreads();
            {
              float tot[kPdGateRowsPerWarp];
              pd_dot_rows<2 * kPredH, kPdGateRowsPerWarp>(s_w, 2 * kPredH, warp, kPdWarps, nunits * 4, s_x, 2 * kPredH, lane, ne, tot);
              if ((lane & 3) == 0) {
                const int my_e = (lane >> 2) & 7;
#pragma unroll
                for (int i = 0; i < kPdGateRowsPerWarp; ++i) {
                  const int r = warp + i * kPdWarps;
                  if (r < nunits * 4) s_gates[r][my_e] = tot[i] + __ldg(bias + (r & 3) * kPredH + u0 + (r >> 2));
                }
              }
            }
            __syncthreads();
            if (tid < nunits * kPdChunk) {
              const int u = tid / kPdChunk, e = tid % kPdChunk;
              if (s_flag[e]) {
                const int ge = e0 + e, j = u0 + u, slot = d.slot[ge];
                float* h = d.pred_h + ((size_t)slot * kPredL + layer) * kPredH;
                float* c = d.pred_c + ((size_t)slot * kPredL + layer) * kPredH;
                const float ig = sigmoidf_acc(s_gates[u * 4 + 0][e]);
                const float fg = sigmoidf_acc(s_gates[u * 4 + 1][e]);
                const float gg = tanhf(s_gates[u * 4 + 2][e]);
                const float og = sigmoidf_acc(s_gates[u * 4 + 3][e]);
                const float cn = fg * __ldcg(c + j) + ig * gg;
                const float hn = og * tanhf(cn);
                c[j] = cn;
                h[j] = hn;
                if (layer == 0) {
                  a.x1[(size_t)ge * 2 * kPredH + j] = hn;
                  a.x1[(size_t)ge * 2 * kPredH + kPredH + j] = __ldcg(d.pred_h + ((size_t)slot * kPredL + 1) * kPredH + j);
                } else {
                  d.pred_g[(size_t)slot * kPredH + j] = hn;
                  a.gvec[(size_t)ge * kPredH + j] = hn;
                }
              }
            }
          }
          __syncthreads();
        }
        if (!pd_grid_sync(a.bar, ++nbar * G, a.err)) return;
      }
      // ================= phase 5: joint.pred rows [u0, u0 + nunits) -> P[slot]
      for (int e0 = 0; e0 < d.B && nunits > 0; e0 += kPdChunk) {
        const int ne = min(kPdChunk, d.B - e0);
        if (tid < kPdChunk) s_flag[tid] = (tid < ne && __ldcg(d.emit_tok + e0 + tid) >= 0) ? 1 : 0;
        __syncthreads();
        bool any = false;
#pragma unroll
        for (int e = 0; e < kPdChunk; ++e) any |= s_flag[e] != 0;
        if (any) {
          for (int i = tid; i < kPdChunk * (kPredH / 4); i += kPdThreads) {
            const int e = i / (kPredH / 4), c = i % (kPredH / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s_flag[e]) v = __ldcg(reinterpret_cast<const float4*>(a.gvec + (size_t)(e0 + e) * kPredH) + c);
            reinterpret_cast<float4*>(s_x + (size_t)e * kPredH)[c] = v;
          }
          __syncthreads();
          {
            float tot[1];
            pd_dot_rows<kPredH, 1>(s_wjp, kPredH, warp, kPdWarps, nunits, s_x, kPredH, lane, ne, tot);
            const int my_e = (lane >> 2) & 7;
            if ((lane & 3) == 0 && warp < nunits && s_flag[my_e])
              d.pred_proj[(size_t)d.slot[e0 + my_e] * kJointH + u0 + warp] = tot[0] + __ldg(a.b_jp + u0 + warp);
          }
        }
        __syncthreads();
      }
      if (!pd_grid_sync(a.bar, ++nbar * G, a.err)) return;
    }
    if (n_active == 0) { ++pass; break; }
  }
  if (cta == 0 && tid == 0) { *a.passes_out = pass; *d.n_active = 0; *d.m_pred = 0; *d.m_joint = 0; }
}

size_t decode_persistent_smem() { return kPdSmem; }

// Cooperative launch (all CTAs co-resident: the grid barrier needs it).  Returns false when the device cannot hold `grid` CTAs of
// this kernel at once (the caller then takes the launch-chain path).
bool launch_decode_persistent(const DecPersistArgs& a, int grid, cudaStream_t st) {
  static int max_ctas = -1;
  if (max_ctas < 0) {
    PKB_CUDA(cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPdSmem));
    int per_sm = 0, dev = 0, sms = 0;
    PKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persistent_kernel, kPdThreads, kPdSmem));
    PKB_CUDA(cudaGetDevice(&dev));
    PKB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    max_ctas = per_sm * sms;
  }
  const int cols_per = (kJointOut + grid - 1) / grid, units_per = (kPredH + grid - 1) / grid;
  if (grid > max_ctas || cols_per > kPdColsMax || units_per > kPdUnitsMax || a.d.B > kPersistMaxB || a.d.B > grid) return false;
  DecPersistArgs args = a;
  void* params[] = {&args};
  PKB_CUDA(cudaLaunchCooperativeKernel((const void*)decode_persistent_kernel, dim3(grid), dim3(kPdThreads), params, kPdSmem, st));
  return true;
}

}  // namespace pkb
