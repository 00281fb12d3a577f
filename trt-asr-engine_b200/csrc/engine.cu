// engine.cu -- see engine.h.  Host orchestration of the batched streaming step.
#include "engine.h"
#include "offline_long.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>

namespace pkb {

namespace {

// Allocation registry: everything allocated while an Engine is being constructed belongs to that Engine and is released by
// its destructor (a process may create and destroy many sessions: parakeet_create_session / parakeet_destroy_session).
// Transient buffers of individual calls are allocated outside a scope and freed by their owners.
thread_local std::vector<void*>* g_dev_scope = nullptr;
thread_local std::vector<void*>* g_host_scope = nullptr;
void* dev_alloc_bytes(size_t bytes) {
  void* p = nullptr;
  PKB_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
  if (g_dev_scope) g_dev_scope->push_back(p);
  return p;
}
template <typename T>
T* dev_alloc(size_t n) { return static_cast<T*>(dev_alloc_bytes(n * sizeof(T))); }
void dev_free(void* p) {      // free a buffer that may have been registered in the current scope
  if (g_dev_scope) {
    auto it = std::find(g_dev_scope->begin(), g_dev_scope->end(), p);
    if (it != g_dev_scope->end()) g_dev_scope->erase(it);
  }
  cudaFree(p);
}
template <typename T>
void host_alloc(T** p, size_t bytes) {
  PKB_CUDA(cudaMallocHost(p, bytes));
  if (g_host_scope) g_host_scope->push_back(*p);
}
template <typename T>
T* dev_upload(const std::vector<T>& v) {
  T* p = dev_alloc<T>(v.size());
  if (!v.empty()) {
    // cudaMemcpy from pageable memory returns once the data is STAGED; the DMA may still be in flight, and the engine's
    // non-blocking stream is not ordered against it -- so wait for the device before anyone can consume the buffer.
    PKB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    PKB_CUDA(cudaDeviceSynchronize());
  }
  return p;
}

// transient device buffers of one call: released on every exit path (exceptions included)
struct DevTmp {
  std::vector<void*> ptrs;
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    PKB_CUDA(cudaMalloc(&p, std::max<size_t>(n * sizeof(T), 1)));
    ptrs.push_back(p);
    return static_cast<T*>(p);
  }
  template <typename T>
  T* upload(const T* host, size_t n) {
    T* p = alloc<T>(n);
    if (n) { PKB_CUDA(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice)); PKB_CUDA(cudaDeviceSynchronize()); }
    return p;
  }
  ~DevTmp() { for (void* p : ptrs) cudaFree(p); }
};

inline int sub_len(int L) { return L <= 0 ? 0 : (L - 1) / 2 + 1; }   // floor((L + 2 - 3)/2) + 1, calc_length of dw_striding

struct GemmW {
  __nv_bfloat16* w = nullptr;
  int N = 0, K = 0;
  TensorMap map;        // 128-row boxes
  TensorMap map32;      // 32-row boxes (tail slices of the CTA-pair kernel)
};

struct ActBuf {
  __nv_bfloat16* ptr = nullptr;
  int rows_cap = 0, K = 0;
  long long lo_off = 0;
  TensorMap map;
  ActOut out() const { return ActOut{ptr, K, lo_off}; }
};

struct LayerW {
  float *n_ff1_g, *n_ff1_b, *n_att_g, *n_att_b, *n_conv_g, *n_conv_b, *n_ff2_g, *n_ff2_b, *n_out_g, *n_out_b;
  GemmW ff1_1, ff1_2, qkv, kv, out, pw1, pw2, ff2_1, ff2_2;
  float *dw_w, *dw_wt, *dw_b, *bias_u, *bias_v;      // dw_wt: dw_w transposed to [9][1024]
  void* ppos_t;                 // precise mode: transposed table [head][128][kPosRows] f32
  __nv_bfloat16* ppos_n;        // bf16 mode: natural table [head][kPosRowsPad][128]
  TensorMap ppos_map;           //           TMA map over it as the W operand [8 * kPosRowsPad, 128] of the position-score GEMM
};

// ---- small utility kernels ----
__global__ void f32_to_act_kernel(const float* __restrict__ src, long long src_row_stride, long long src_col_stride, int rows,
                                  int cols, ActOut a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  store_act(a.ptr, r, a.lda, c, src[r * src_row_stride + c * src_col_stride], a.lo_off);
}
// hidden[(b*T + t)*U + u] = relu(E[b*T+t] + P[b*U+u])
__global__ void joint_hidden_grid_kernel(const float* __restrict__ E, const float* __restrict__ P, int T, int U, ActOut a) {
  const int row = blockIdx.x;
  const int u = row % U, bt = row / U, b = bt / T;
  for (int c = threadIdx.x; c < kJointH; c += blockDim.x)
    store_act(a.ptr, row, a.lda, c, fmaxf(E[(size_t)bt * kJointH + c] + P[((size_t)b * U + u) * kJointH + c], 0.0f), a.lo_off);
}
// P_n[h][r][d] <- P[r][h*128+d]   (bf16, rows >= kPosRows stay zero)
__global__ void ppos_natural_kernel(const float* __restrict__ P, __nv_bfloat16* __restrict__ out) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < kDModel; c += blockDim.x)
    out[((size_t)(c >> 7) * kPosRowsPad + r) * kDHead + (c & 127)] = __float2bfloat16_rn(P[(size_t)r * kDModel + c]);
}
// P^T[h][d][r] <- P[r][h*128+d]
__global__ void ppos_transpose_kernel(const float* __restrict__ P, void* __restrict__ out, int is_f32) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < kDModel; c += blockDim.x) {
    const size_t o = (size_t)c * kPosRows + r;   // c == h*128+d
    const float v = P[(size_t)r * kDModel + c];
    if (is_f32) ((float*)out)[o] = v; else ((__nv_bfloat16*)out)[o] = __float2bfloat16_rn(v);
  }
}
__global__ void fill_import_rows_kernel(int* row_entry, int* row_pos, int n_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  row_entry[i] = i / kCacheS;
  row_pos[i] = (i % kCacheS) - kCacheS;   // logical cache position j -> ring index head + 256 + (j - 256)
}

// dst[slot][off + k] = src[i*stride + k]
__global__ void audio_append_kernel(const float* __restrict__ src, long long stride, float* __restrict__ audio_buf,
                                    const int* __restrict__ slot, const int* __restrict__ off, const int* __restrict__ cnt) {
  pdl_enter();
  const int i = blockIdx.y;
  const int n = cnt[i];
  float* d = audio_buf + (size_t)slot[i] * 32768 + off[i];
  const float* s = src + (size_t)i * stride;
  for (int k = blockIdx.x * 1024 + threadIdx.x; k < min(n, (int)(blockIdx.x + 1) * 1024); k += 256) d[k] = s[k];
}
// phase 0: tmp[slot][k] = buf[slot][off + k];  phase 1: buf[slot][k] = tmp[slot][k]   (k < fill)
__global__ void audio_move_kernel(float* __restrict__ buf, float* __restrict__ tmp, const int* __restrict__ slot,
                                  const int* __restrict__ off, const int* __restrict__ fill, int phase) {
  const int i = blockIdx.y;
  const size_t base = (size_t)slot[i] * 32768;
  const int n = fill[i], o = off[i];
  for (int k = blockIdx.x * 1024 + threadIdx.x; k < min(n, (int)(blockIdx.x + 1) * 1024); k += 256) {
    if (phase == 0) tmp[base + k] = buf[base + o + k]; else buf[base + k] = tmp[base + k];
  }
}

// NaN / Inf census of up to 4 row segments (f32 or bf16): out = {nan count, inf count, first nan index, first inf index}
struct GuardRows { long long off[4]; int n_rows; int row_len; };
__global__ void guard_scan_kernel(const void* __restrict__ base, int is_bf16, GuardRows rows, int* __restrict__ out) {
  const int total = rows.n_rows * rows.row_len;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / rows.row_len, c = i - r * rows.row_len;
    const float v = is_bf16 ? __bfloat162float(((const __nv_bfloat16*)base)[rows.off[r] + c]) : ((const float*)base)[rows.off[r] + c];
    if (v != v) { atomicAdd(out, 1); atomicMin(out + 2, i); }
    else if (isinf(v)) { atomicAdd(out + 1, 1); atomicMin(out + 3, i); }
  }
}

}  // namespace

// ================================================================================================
struct Engine::Stream {
  bool open = false;
  int slot = 0;
  long long frames_written = 0;          // frames ever written to the feature ring
  std::deque<Entry> pending;             // explicit chunks (legacy ABI granularity)
  bool audio_mode = false;
  bool offline = false;                  // offline (full-context, cache-free) encoding of every push
  std::vector<float> audio;              // host overflow FIFO: samples that did not fit the device audio buffer yet
  bool needs_prime = false;              // predictor priming is deferred to the next batched pass (one launch set for all)
  int dev_off = 0;                       // first valid sample inside this stream's device audio buffer (always even)
  int dev_fill = 0;                      // valid samples from dev_off
  long long sched_chunk = 0;             // index of the next scheduled chunk (audio mode)
  bool has_norm = false;
  bool run_norm = false;                 // streaming-safe running mean / std normalisation of the GPU frontend's frames
  int cache_len = 0;
  int head = 0;
  long long chunks = 0;
  std::vector<int> tokens;
  std::vector<int> token_frames;         // encoder frame (80 ms timebase) each token was emitted on, counted from the utterance start
  long long enc_frames = 0;              // encoder frames decoded so far (the live edge of the transcript)
  long long lf_store_off = -1;           // >= 0: the utterance's joint encoder projections wait in the deferred-decode store at this row
  int lf_store_T = 0;
  int last_entry = -1, last_out_T = 0;   // where the stream's encoder_output of the last batched pass sits in enc_out [B,1024,out_T]
  ChunkResult last;
};

// one captured step shape (see Engine::step_graph)
struct Engine::StepGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long fixed_launches = 0;      // kernels outside the loop
  int body_launches = 0;             // kernels per loop pass
  cudaEvent_t ev_loop0 = nullptr, ev_loop1 = nullptr;      // recorded by the graph right before / after the WHILE node
};

constexpr int kRunStats = 1 + 2 * kNMels + 3;      // padded to a multiple of 4 floats
constexpr int kProfClasses = 5;   // 0 tcgen05 GEMM, 1 streaming attention, 2 log-mel, 3 decode loop, 4 whole-utterance attention

struct Engine::Impl {
  std::vector<void*> dev_allocs, host_allocs;      // owned device / pinned-host buffers (see g_dev_scope)
  Frontend frontend;
  // weights
  SubsampleWeights sub{};
  GemmW sub_pw1, sub_pw2, sub_out;
  float *sub_pw1_b = nullptr, *sub_pw2_b = nullptr, *sub_out_b = nullptr;
  std::vector<LayerW> layers;
  GemmW joint_enc, joint_pred, joint_out, lstm[2];
  float *joint_enc_b = nullptr, *joint_pred_b = nullptr, *joint_out_b = nullptr, *lstm_b[2] = {nullptr, nullptr};
  __nv_bfloat16* embed = nullptr;
  unsigned* punct_bits = nullptr;
  // per-slot state
  void *kring = nullptr, *vring = nullptr, *acache = nullptr;
  bool attn_mma = false;                 // bf16 mode: natural-layout K ring + tensor-core attention (attn_mma.cu)
  TensorMap map_k, map_v;                // TMA maps over the K / V rings of all layers
  TensorMap map_k32, map_v32, map_k8, map_v8;   // ... with 32-row / 8-row boxes (partly valid blocks: valid 8-key groups only)
  size_t ring_layer_elems = 0;           // elements per layer in each ring (slots * 288 * 1024)
  float* cache_tm = nullptr;             // [slots][L][1024][4]
  float* feat_ring = nullptr;            // [slots][kFeatRing][128]
  float* norm_stats = nullptr;           // [slots][2][128]
  float* run_stats = nullptr;            // [slots][kRunStats]: frame count, running mean[128], M2[128] (running normalisation)
  float *pred_h = nullptr, *pred_c = nullptr, *pred_g = nullptr, *pred_proj = nullptr;
  int *n_emitted = nullptr, *y_id = nullptr;
  // work buffers
  int Mcap = 0, Bcap = 0, T3cap = 0, T2cap = 0;
  ActBuf a_y1, a_sub1, a_sub2, a_sub3, a_ln, a_ff, a_xf, a_hid, a_pred, a_g, a_imp, a_pos;
  __nv_bfloat16* q_bf16 = nullptr;       // bf16 mode: [2][q_rows,1024] (q + pos_bias_u | q + pos_bias_v), q_rows = Mcap padded to 128
  long long q_plane = 0;                 //           elements between the planes
  TensorMap map_qv;                      //           TMA map over the q + pos_bias_v plane (A operand of the position-score GEMM)
  __nv_bfloat16* g_pos = nullptr;        //           position scores [q_rows][8][kPosRowsPad] bf16
  float *x = nullptr, *q = nullptr, *cglu = nullptr, *enc_proj = nullptr, *logits = nullptr, *gates = nullptr,
        *enc_out = nullptr, *ppos_tmp = nullptr, *scratch_f32 = nullptr, *part_val = nullptr, *dur_logits = nullptr;
  int* part_idx = nullptr;
  float* part_ws = nullptr;              // split-K partial sums [4][part_rows][1024] (deferred residual)
  int part_rows = 0, part_rows_split = 0;
  size_t scratch_f32_elems = 0;
  int* batch_ints = nullptr;             // device: entry arrays + prefixes
  int* batch_ints_host = nullptr;        // pinned
  int *row_entry = nullptr, *row_pos = nullptr, *rowmap3 = nullptr;
  int *imp_row_entry = nullptr, *imp_row_pos = nullptr;
  // decode buffers
  int *t_cur = nullptr, *n_sym = nullptr, *active = nullptr, *emit_tok = nullptr, *pred_rowmap = nullptr, *n_steps = nullptr,
      *steps = nullptr, *counters = nullptr, *force_toks = nullptr;
  int* res_host = nullptr;               // pinned: [Bcap] n_steps + [Bcap*32*3] steps
  int* counters_host = nullptr;          // pinned [4]: n_active of the last two decode iterations at [0] and [2]
  cudaEvent_t dec_events[2] = {nullptr, nullptr};
  // frontend staging
  float* audio_buf = nullptr;            // [slots][kAudioCap] per-stream device audio (carry + not yet framed samples)
  float* audio_tmp = nullptr;            // [slots][kAudioCap] compaction scratch
  float* audio_stage = nullptr;          // device staging for batched pushes from host memory
  float* audio_stage_host = nullptr;     // pinned bounce buffer for pageable sources
  size_t audio_stage_cap = 0;
  int* push_meta = nullptr;              // device [3][Bcap]: slot, dst offset, count
  int* push_meta_host = nullptr;         // pinned
  std::vector<cudaEvent_t> user_events;
  // optional per-launch timing of the tensor-core GEMM (bench roofline)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  std::vector<double> prof_flops;        // algorithmic work of the bracketed launch: FLOPs (class 0) or bytes (classes 1, 2)
  std::vector<int> prof_class;           // 0 = tcgen05 GEMM, 1 = attention, 2 = log-mel frontend
  size_t prof_used = 0;
  double prof_ms[kProfClasses] = {}, prof_work[kProfClasses] = {};
  long long prof_launches[kProfClasses] = {};
  // whole-utterance offline path (allocated on first use; see Engine::offline_utterances)
  std::vector<GemmW> lf_wpos;            // linear_pos weights per layer (the streaming path only keeps its 320-row projected table)
  float* lf_feat = nullptr;              // [lf_feat_frames][128] normalised log-mel frames of the utterances of one call
  size_t lf_feat_frames = 0;
  ActBuf lf_pos_act;                     // sinusoid table operand [2*Mcap][1024] (hi + lo planes)
  void* lf_qkv = nullptr;                // [Mcap][3072] q | k | v  (bf16, f32 in precise mode)
  void* lf_ppos = nullptr;               // [2*Mcap][1024] this layer's projected table (bf16 / f32)
  int* lf_steps = nullptr;               // decode trace [B][lf_steps_per][3]
  float* lf_store = nullptr;             // deferred decode: joint.enc(encoder rows) of encoded, not yet decoded utterances [rows][640]
  size_t lf_store_cap = 0, lf_store_used = 0;
  TensorMap lf_map_qkv, lf_map_pos;      // TMA maps of this call's q|k|v rows and projected table (bf16 mode)
  // tcgen05 whole-utterance attention (lf_attn_tc.cu): biased query planes, transposed V and their maps
  __nv_bfloat16* lf_qplanes = nullptr;   // [2][lf_q_rows][1024]
  int lf_q_rows = 0;
  __nv_bfloat16* lf_vt = nullptr;        // [1024][lf_ldv]
  long long lf_ldv = 0;
  TensorMap lf_map_q, lf_map_pos64, lf_map_vt;
  int lf_attn_kind = 2;                  // PARAKEET_B200_LF_ATTN: 2 = tcgen05 (default), 1 = mma.sync fed by TMA, 0 = mma.sync fed by LDG/STS
  size_t lf_steps_ints = 0;
  FrontSegment* segs_dev = nullptr;
  FrontSegment* segs_host = nullptr;
  int* fprefix_dev = nullptr;
  int* fprefix_host = nullptr;
  // feature pushes (the legacy ABI's input): chunks queued since the last batched pass wait bins-major in one pinned staging
  // area; ONE H2D copy + ONE transpose kernel per pass move all of them into the per-stream feature rings
  float* fstage_dev = nullptr;
  float* fstage_host = nullptr;
  size_t fstage_cap = 0, fstage_used = 0;      // floats
  FeatPush* fpush_dev = nullptr;
  FeatPush* fpush_host = nullptr;
  int fpush_cap = 0, fpush_n = 0, fpush_max_T = 0;
  bool fstage_inflight = false;                // flushed to the device, host area not yet known to be consumed
  bool frontend_inflight = false;              // frontend_pass() queued copies from its pinned descriptors, no synchronisation since
  // whole-step CUDA graphs, one per step shape (see Engine::step_graph)
  std::map<std::array<int, 6>, std::unique_ptr<Engine::StepGraph>> graphs;
  std::map<std::array<int, 6>, int> graph_seen;
  int graph_mode = 1, graph_pdl = 1;
  cudaGraph_t lf_loop_graph = nullptr;         // decode-loop graph of the whole-utterance call in flight (run_decode_loop_graph)
  cudaGraphExec_t lf_loop_exec = nullptr;
  int lf_loop_body_launches = 0, lf_loop_prof_idx = -1, lf_loop_B = 0;
  // persistent decode loop (dec_kernels.cu): scratch for <= kPersistMaxB entries, allocated on first use
  bool lf_persist_active = false;
  unsigned* pd_bar = nullptr;      // [0] barrier counter, [1] error flag, [2..5] per-pass flags, [6] passes
  float *pd_part_val = nullptr, *pd_dur = nullptr, *pd_xin = nullptr, *pd_x1 = nullptr, *pd_gvec = nullptr;
  int* pd_part_idx = nullptr;
  double loop_ms = 0.0, loop_bytes = 0.0;      // decode loops run inside step graphs since the last decode_loop_stats(reset)
  long long loop_passes = 0, loop_count = 0;
  int* meta2 = nullptr;                        // device [2]: (slot, head) of a single-stream import / export
  int* guard_dev = nullptr;                    // device [4]: nan count, inf count, first nan index, first inf index (NaN guard)
  int* guard_host = nullptr;                   // pinned
};

constexpr int kFeatRing = 512;           // frames kept per stream (5.12 s)
constexpr int kAudioCap = 32768;         // samples per stream resident on the device (2 s)
constexpr int kNumBatchFields = 13;

// ================================================================================================
Engine::Engine(const EngineOptions& opt) : opt_(opt) {
  PKB_CHECK(opt_.max_streams >= 1, "max_streams must be >= 1");
  PKB_CUDA(cudaSetDevice(opt_.device_id));
  cudaDeviceProp prop;
  PKB_CUDA(cudaGetDeviceProperties(&prop, opt_.device_id));
  sm_count_ = prop.multiProcessorCount;
  PKB_CHECK(prop.major == 10, "this library is built for sm_100a (B200) only; found compute capability " +
                                  std::to_string(prop.major) + "." + std::to_string(prop.minor));
  PKB_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
  im_.reset(new Impl());
  struct ScopeGuard {
    ScopeGuard(std::vector<void*>* d, std::vector<void*>* h) { g_dev_scope = d; g_host_scope = h; }
    ~ScopeGuard() { g_dev_scope = nullptr; g_host_scope = nullptr; }
  } scope(&im_->dev_allocs, &im_->host_allocs);
  vocab_ = Vocab(opt_.model_dir + "/vocab.txt");
  PKB_CHECK(vocab_.size() <= kVocab, "vocab.txt must have 1..8193 lines");
  tok_start_ = vocab_.find("<|startoftranscript|>");
  tok_lang_ = vocab_.find("<|en|>");
  punct_bits_ = vocab_.punct_bitmap(kVocab);
  { const char* v = getenv("PARAKEET_B200_GRAPH"); if (v) im_->graph_mode = atoi(v); }
  { const char* v = getenv("PARAKEET_B200_GRAPH_PDL"); if (v) im_->graph_pdl = atoi(v); }
  load_weights();
  alloc_state();
  streams_.resize(opt_.max_streams);
  for (int i = 0; i < opt_.max_streams; ++i) {
    streams_[i].reset(new Stream());
    streams_[i]->slot = i;
  }
  PKB_CUDA(cudaStreamSynchronize(st_));
  PKB_CUDA(cudaDeviceSynchronize());   // init-time uploads went through the legacy stream (staged pageable copies)
}

Engine::~Engine() {
  if (st_) cudaStreamSynchronize(st_);
  if (im_) {
    for (auto& kv : im_->graphs) {
      cudaGraphExecDestroy(kv.second->exec); cudaGraphDestroy(kv.second->graph);
      cudaEventDestroy(kv.second->ev_loop0); cudaEventDestroy(kv.second->ev_loop1);
    }
    if (im_->lf_loop_exec) { cudaGraphExecDestroy(im_->lf_loop_exec); cudaGraphDestroy(im_->lf_loop_graph); }
    for (auto& e : im_->user_events) cudaEventDestroy(e);
    for (auto& e : im_->dec_events) if (e) cudaEventDestroy(e);
    for (auto& pr : im_->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (void* p : im_->dev_allocs) cudaFree(p);
    for (void* p : im_->host_allocs) cudaFreeHost(p);
  }
  if (st_) cudaStreamDestroy(st_);
}

void Engine::synchronize() { PKB_CUDA(cudaStreamSynchronize(st_)); }
Engine::Impl* Engine::impl() { return im_.get(); }

int Engine::event_record() {
  cudaEvent_t e;
  PKB_CUDA(cudaEventCreate(&e));
  PKB_CUDA(cudaEventRecord(e, st_));
  im_->user_events.push_back(e);
  return (int)im_->user_events.size() - 1;
}
double Engine::event_elapsed_ms(int a, int b) {
  PKB_CHECK(a >= 0 && b >= 0 && a < (int)im_->user_events.size() && b < (int)im_->user_events.size(), "bad event id");
  PKB_CUDA(cudaEventSynchronize(im_->user_events[b]));
  float ms = 0;
  PKB_CUDA(cudaEventElapsedTime(&ms, im_->user_events[a], im_->user_events[b]));
  return ms;
}
void Engine::profile_enable(bool on) {
  PKB_CUDA(cudaStreamSynchronize(st_));
  im_->profile = on;
  im_->prof_used = 0;
  for (int c = 0; c < kProfClasses; ++c) { im_->prof_ms[c] = im_->prof_work[c] = 0; im_->prof_launches[c] = 0; }
}
// profile mode: bracket the next launch with two CUDA events on the engine stream
int Engine::prof_begin(int cls, double work) {
  if (im_->profile) {      // launches that are being captured into a graph have no timing of their own (their events would be graph nodes)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st_, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return -1;
  }
  Impl& im = *im_;
  if (!im.profile) return -1;
  if (im.prof_used == im.prof_events.size()) {
    cudaEvent_t a0, a1;
    PKB_CUDA(cudaEventCreate(&a0));
    PKB_CUDA(cudaEventCreate(&a1));
    im.prof_events.emplace_back(a0, a1);
    im.prof_flops.push_back(0.0);
    im.prof_class.push_back(0);
  }
  const int i = (int)im.prof_used++;
  im.prof_flops[i] = work;
  im.prof_class[i] = cls;
  PKB_CUDA(cudaEventRecord(im.prof_events[i].first, st_));
  return i;
}
void Engine::prof_end(int i) {
  if (i >= 0) PKB_CUDA(cudaEventRecord(im_->prof_events[i].second, st_));
}
void Engine::profile_collect() {      // call after a synchronised step
  Impl& im = *im_;
  for (size_t i = 0; i < im.prof_used; ++i) {
    float ms = 0;
    PKB_CUDA(cudaEventElapsedTime(&ms, im.prof_events[i].first, im.prof_events[i].second));
    const int c = im.prof_class[i];
    im.prof_ms[c] += ms;
    im.prof_work[c] += im.prof_flops[i];
    im.prof_launches[c] += 1;
  }
  im.prof_used = 0;
}
// Device time of the decode loops executed inside step graphs (CUDA events recorded by the graph around its WHILE node), their
// passes (= symbols decoded per stream, max over the batch) and algorithmic bytes; reset != 0 clears the counters afterwards.
void Engine::decode_loop_stats(double* ms, double* bytes, long long* passes, long long* loops, int reset) {
  Impl& im = *im_;
  *ms = im.loop_ms; *bytes = im.loop_bytes; *passes = im.loop_passes; *loops = im.loop_count;
  if (reset) { im.loop_ms = im.loop_bytes = 0.0; im.loop_passes = im.loop_count = 0; }
}
int Engine::graphs_built() const { return (int)im_->graphs.size(); }
// PARAKEET_BLANK_PENALTY at run time.  Kernel arguments of captured steps carry the old value: the step graphs are dropped.
void Engine::set_blank_penalty(float p) {
  PKB_CUDA(cudaStreamSynchronize(st_));
  opt_.blank_penalty = p;
  for (auto& kv : im_->graphs) {
    cudaGraphExecDestroy(kv.second->exec); cudaGraphDestroy(kv.second->graph);
    cudaEventDestroy(kv.second->ev_loop0); cudaEventDestroy(kv.second->ev_loop1);
  }
  im_->graphs.clear();
  im_->graph_seen.clear();
}
void Engine::profile_read(int cls, double* ms, double* work, long long* launches) {
  PKB_CHECK(cls >= 0 && cls < kProfClasses, "profile class");
  *ms = im_->prof_ms[cls]; *work = im_->prof_work[cls]; *launches = im_->prof_launches[cls];
}

// ------------------------------------------------------------------------------------------------ weights
static GemmW upload_gemm_w(const std::vector<uint16_t>& bits, int N, int K) {
  GemmW g;
  g.N = N;
  g.K = K;
  PKB_CHECK((size_t)N * K == bits.size(), "weight shape mismatch");
  // pad rows to a multiple of 128 so that every 128-row TMA box lies inside the allocation
  const size_t rows_pad = ((size_t)N + 127) / 128 * 128;
  g.w = dev_alloc<__nv_bfloat16>(rows_pad * K);
  PKB_CUDA(cudaMemset(g.w, 0, rows_pad * K * 2));
  PKB_CUDA(cudaMemcpy(g.w, bits.data(), bits.size() * 2, cudaMemcpyHostToDevice));   // synchronous: both are complete on return
  PKB_CUDA(cudaDeviceSynchronize());
  make_tensor_map_2d(&g.map, g.w, rows_pad, K, K, 128);
  make_tensor_map_2d(&g.map32, g.w, rows_pad, K, K, 32);
  return g;
}

// NOTE: the engine stream is non-blocking, so legacy-default-stream memsets would NOT be ordered against it; every
// device-side initialisation below is therefore issued on the engine stream itself.
static ActBuf make_act(int rows_cap, int K, bool split, cudaStream_t st) {
  ActBuf a;
  a.rows_cap = (rows_cap + 127) / 128 * 128;
  a.K = K;
  const size_t plane = (size_t)a.rows_cap * K;
  a.ptr = dev_alloc<__nv_bfloat16>(plane * (split ? 2 : 1));
  PKB_CUDA(cudaMemsetAsync(a.ptr, 0, plane * (split ? 2 : 1) * 2, st));
  a.lo_off = split ? (long long)plane : 0;
  make_tensor_map_2d(&a.map, a.ptr, (uint64_t)a.rows_cap * (split ? 2 : 1), K, K, 128);
  return a;
}

void Engine::load_weights() {
  WeightsFile wf(opt_.model_dir + "/weights.bin");
  L_ = (int)wf.cfg("n_layers");
  PKB_CHECK(wf.cfg("d_model") == kDModel && wf.cfg("n_heads") == kHeads && wf.cfg("ff_dim") == kFF &&
                wf.cfg("conv_kernel") == kConvK && wf.cfg("sub_channels") == kSubCh && wf.cfg("feat_in") == kNMels &&
                wf.cfg("vocab") == kVocab && wf.cfg("n_dur") == kNDur && wf.cfg("pred_hidden") == kPredH &&
                wf.cfg("pred_layers") == kPredL && wf.cfg("joint_hidden") == kJointH && wf.cfg("cache_size") == kCacheS &&
                wf.cfg("time_ctx") == kTimeCtx && wf.cfg("cache_drop") == kCacheDrop && wf.cfg("valid_out_len") == kValidOut &&
                wf.cfg("drop_extra_pre_encoded") == kDropPre,
            "weights.bin architecture constants do not match this build (Parakeet-TDT-0.6B-v3 shapes are compiled in)");
  Impl& im = *im_;
  // every vector-shaped tensor is checked against the element count the kernels will read
  auto F = [&](const std::string& name, size_t n) {
    std::vector<float> v = wf.f32(name);
    PKB_CHECK(v.size() == n, "weights.bin: tensor " + name + " has " + std::to_string(v.size()) + " elements, expected " + std::to_string(n));
    return v;
  };
  const std::string pe = "encoder.pre_encode.";
  im.sub.w0 = dev_upload(F(pe + "conv.0.weight", (size_t)kSubCh * 9));
  im.sub.b0 = dev_upload(F(pe + "conv.0.bias", kSubCh));
  im.sub.w2 = dev_upload(F(pe + "conv.2.weight", (size_t)kSubCh * 9));
  im.sub.b2 = dev_upload(F(pe + "conv.2.bias", kSubCh));
  im.sub.w5 = dev_upload(F(pe + "conv.5.weight", (size_t)kSubCh * 9));
  im.sub.b5 = dev_upload(F(pe + "conv.5.bias", kSubCh));
  {
    auto transposed = [&](const std::string& name) {
      const std::vector<float> w = F(name, (size_t)kSubCh * 9);
      std::vector<float> t((size_t)9 * kSubCh);
      for (int c = 0; c < kSubCh; ++c)
        for (int k = 0; k < 9; ++k) t[(size_t)k * kSubCh + c] = w[(size_t)c * 9 + k];
      return dev_upload(t);
    };
    im.sub.w0t = transposed(pe + "conv.0.weight");
    im.sub.w2t = transposed(pe + "conv.2.weight");
    im.sub.w5t = transposed(pe + "conv.5.weight");
  }
  im.sub_pw1 = upload_gemm_w(wf.bf16(pe + "conv.3.weight"), kSubCh, kSubCh);
  im.sub_pw1_b = dev_upload(F(pe + "conv.3.bias", kSubCh));
  im.sub_pw2 = upload_gemm_w(wf.bf16(pe + "conv.6.weight"), kSubCh, kSubCh);
  im.sub_pw2_b = dev_upload(F(pe + "conv.6.bias", kSubCh));
  {
    // Linear(4096 -> 1024): NeMo flattens [C=256, F=16] channel-major (c*16+f); our operand rows are (f*256+c)
    std::vector<uint16_t> w = wf.bf16(pe + "out.weight"), p(w.size());
    for (int n = 0; n < kDModel; ++n)
      for (int c = 0; c < kSubCh; ++c)
        for (int f = 0; f < 16; ++f) p[(size_t)n * 4096 + f * kSubCh + c] = w[(size_t)n * 4096 + c * 16 + f];
    im.sub_out = upload_gemm_w(p, kDModel, 4096);
    im.sub_out_b = dev_upload(F(pe + "out.bias", kDModel));
  }
  im.layers.resize(L_);
  for (int l = 0; l < L_; ++l) {
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    LayerW& w = im.layers[l];
    w.n_ff1_g = dev_upload(F(p + "norm_feed_forward1.weight", kDModel)); w.n_ff1_b = dev_upload(F(p + "norm_feed_forward1.bias", kDModel));
    w.n_att_g = dev_upload(F(p + "norm_self_att.weight", kDModel));      w.n_att_b = dev_upload(F(p + "norm_self_att.bias", kDModel));
    w.n_conv_g = dev_upload(F(p + "norm_conv.weight", kDModel));         w.n_conv_b = dev_upload(F(p + "norm_conv.bias", kDModel));
    w.n_ff2_g = dev_upload(F(p + "norm_feed_forward2.weight", kDModel)); w.n_ff2_b = dev_upload(F(p + "norm_feed_forward2.bias", kDModel));
    w.n_out_g = dev_upload(F(p + "norm_out.weight", kDModel));           w.n_out_b = dev_upload(F(p + "norm_out.bias", kDModel));
    w.ff1_1 = upload_gemm_w(wf.bf16(p + "feed_forward1.linear1.weight"), kFF, kDModel);
    w.ff1_2 = upload_gemm_w(wf.bf16(p + "feed_forward1.linear2.weight"), kDModel, kFF);
    w.ff2_1 = upload_gemm_w(wf.bf16(p + "feed_forward2.linear1.weight"), kFF, kDModel);
    w.ff2_2 = upload_gemm_w(wf.bf16(p + "feed_forward2.linear2.weight"), kDModel, kFF);
    {
      std::vector<uint16_t> qkv = wf.bf16(p + "self_attn.linear_q.weight");
      std::vector<uint16_t> k = wf.bf16(p + "self_attn.linear_k.weight"), v = wf.bf16(p + "self_attn.linear_v.weight");
      qkv.insert(qkv.end(), k.begin(), k.end());
      qkv.insert(qkv.end(), v.begin(), v.end());
      w.qkv = upload_gemm_w(qkv, 3 * kDModel, kDModel);
      w.kv.w = w.qkv.w + (size_t)kDModel * kDModel;   // rows [1024,3072): re-projection of imported caches
      w.kv.N = 2 * kDModel;
      w.kv.K = kDModel;
      make_tensor_map_2d(&w.kv.map, w.kv.w, 2 * kDModel, kDModel, kDModel, 128);
    }
    w.out = upload_gemm_w(wf.bf16(p + "self_attn.linear_out.weight"), kDModel, kDModel);
    {
      // pointwise_conv1 [2048,1024,1]: interleave (value j, gate j) rows so the GLU pairs adjacent accumulator columns
      std::vector<uint16_t> s = wf.bf16(p + "conv.pointwise_conv1.weight"), d(s.size());
      for (int j = 0; j < kDModel; ++j) {
        memcpy(&d[(size_t)(2 * j) * kDModel], &s[(size_t)j * kDModel], kDModel * 2);
        memcpy(&d[(size_t)(2 * j + 1) * kDModel], &s[(size_t)(j + kDModel) * kDModel], kDModel * 2);
      }
      w.pw1 = upload_gemm_w(d, 2 * kDModel, kDModel);
    }
    w.pw2 = upload_gemm_w(wf.bf16(p + "conv.pointwise_conv2.weight"), kDModel, kDModel);
    {
      // fold eval-mode BatchNorm1d into the depthwise kernel: y = dw*s + (beta - mean*s), s = gamma/sqrt(var+eps)
      std::vector<float> dw = F(p + "conv.depthwise_conv.weight", (size_t)kDModel * kConvK), g = F(p + "conv.batch_norm.weight", kDModel),
                         b = F(p + "conv.batch_norm.bias", kDModel), mu = F(p + "conv.batch_norm.running_mean", kDModel),
                         var = F(p + "conv.batch_norm.running_var", kDModel);
      std::vector<float> off(kDModel);
      for (int c = 0; c < kDModel; ++c) {
        const float s = g[c] / sqrtf(var[c] + 1e-5f);
        for (int k = 0; k < kConvK; ++k) dw[(size_t)c * kConvK + k] *= s;
        off[c] = b[c] - mu[c] * s;
      }
      w.dw_w = dev_upload(dw);
      w.dw_b = dev_upload(off);
      std::vector<float> dwt((size_t)kConvK * kDModel);
      for (int c = 0; c < kDModel; ++c)
        for (int k = 0; k < kConvK; ++k) dwt[(size_t)k * kDModel + c] = dw[(size_t)c * kConvK + k];
      w.dw_wt = dev_upload(dwt);
    }
    w.bias_u = dev_upload(F(p + "self_attn.pos_bias_u", kDModel));
    w.bias_v = dev_upload(F(p + "self_attn.pos_bias_v", kDModel));
    w.ppos_t = nullptr;   // filled in alloc_state (needs work buffers)
    w.ppos_n = nullptr;
  }
  // predictor
  {
    std::vector<uint16_t> emb = wf.bf16("decoder.prediction.embed.weight");
    PKB_CHECK(emb.size() == (size_t)kVocab * kPredH, "embedding shape");
    im.embed = reinterpret_cast<__nv_bfloat16*>(dev_upload(emb));
    for (int l = 0; l < kPredL; ++l) {
      const std::string p = "decoder.prediction.dec_rnn.lstm.";
      std::vector<uint16_t> ih = wf.bf16(p + "weight_ih_l" + std::to_string(l)), hh = wf.bf16(p + "weight_hh_l" + std::to_string(l));
      std::vector<uint16_t> cat((size_t)4 * kPredH * 2 * kPredH);
      for (int n = 0; n < 4 * kPredH; ++n) {
        memcpy(&cat[(size_t)n * 2 * kPredH], &ih[(size_t)n * kPredH], kPredH * 2);
        memcpy(&cat[(size_t)n * 2 * kPredH + kPredH], &hh[(size_t)n * kPredH], kPredH * 2);
      }
      im.lstm[l] = upload_gemm_w(cat, 4 * kPredH, 2 * kPredH);
      std::vector<float> bi = F(p + "bias_ih_l" + std::to_string(l), 4 * kPredH), bh = F(p + "bias_hh_l" + std::to_string(l), 4 * kPredH);
      for (size_t i = 0; i < bi.size(); ++i) bi[i] += bh[i];
      im.lstm_b[l] = dev_upload(bi);
    }
  }
  im.joint_enc = upload_gemm_w(wf.bf16("joint.enc.weight"), kJointH, kDModel);
  im.joint_enc_b = dev_upload(F("joint.enc.bias", kJointH));
  im.joint_pred = upload_gemm_w(wf.bf16("joint.pred.weight"), kJointH, kPredH);
  im.joint_pred_b = dev_upload(F("joint.pred.bias", kJointH));
  im.joint_out = upload_gemm_w(wf.bf16("joint.joint_net.2.weight"), kJointOut, kJointH);
  im.joint_out_b = dev_upload(F("joint.joint_net.2.bias", kJointOut));
  im.punct_bits = dev_upload(punct_bits_);
  // keep the (host) linear_pos weights for alloc_state via a second open: cheap, mmap
}

// ------------------------------------------------------------------------------------------------ state / buffers
void Engine::alloc_state() {
  Impl& im = *im_;
  const bool split = opt_.precision == 1;
  const size_t S = (size_t)opt_.max_streams;
  const size_t kv_elem = split ? 4 : 2;
  im.ring_layer_elems = S * kRingCap * kDModel;
  im.kring = dev_alloc_bytes(im.ring_layer_elems * L_ * kv_elem);
  im.vring = dev_alloc_bytes(im.ring_layer_elems * L_ * kv_elem);
  PKB_CUDA(cudaMemsetAsync(im.kring, 0, im.ring_layer_elems * L_ * kv_elem, st_));
  PKB_CUDA(cudaMemsetAsync(im.vring, 0, im.ring_layer_elems * L_ * kv_elem, st_));
  im.attn_mma = !split;
  if (im.attn_mma) {
    // head-major rings [layer][slot][head][kRingCap][128]: a 2-D map over rows of 128 elements
    make_tensor_map_2d(&im.map_k, im.kring, (uint64_t)L_ * S * kHeads * kRingCap, kDHead, kDHead, 96);
    make_tensor_map_2d(&im.map_v, im.vring, (uint64_t)L_ * S * kHeads * kRingCap, kDHead, kDHead, 96);
    make_tensor_map_2d(&im.map_k32, im.kring, (uint64_t)L_ * S * kHeads * kRingCap, kDHead, kDHead, 32);
    make_tensor_map_2d(&im.map_v32, im.vring, (uint64_t)L_ * S * kHeads * kRingCap, kDHead, kDHead, 32);
    make_tensor_map_2d(&im.map_k8, im.kring, (uint64_t)L_ * S * kHeads * kRingCap, kDHead, kDHead, 8);
    make_tensor_map_2d(&im.map_v8, im.vring, (uint64_t)L_ * S * kHeads * kRingCap, kDHead, kDHead, 8);
  }
  if (opt_.contract_cache) {
    im.acache = dev_alloc_bytes(im.ring_layer_elems * L_ * kv_elem);
    PKB_CUDA(cudaMemsetAsync(im.acache, 0, im.ring_layer_elems * L_ * kv_elem, st_));
  }
  im.cache_tm = dev_alloc<float>(S * L_ * kDModel * kTimeCtx);
  im.feat_ring = dev_alloc<float>(S * kFeatRing * kNMels);
  im.norm_stats = dev_alloc<float>(S * 2 * kNMels);
  im.run_stats = dev_alloc<float>(S * kRunStats);
  PKB_CUDA(cudaMemsetAsync(im.run_stats, 0, S * kRunStats * 4, st_));
  im.pred_h = dev_alloc<float>(S * kPredL * kPredH);
  im.pred_c = dev_alloc<float>(S * kPredL * kPredH);
  im.pred_g = dev_alloc<float>(S * kPredH);
  im.pred_proj = dev_alloc<float>(S * kJointH);
  im.n_emitted = dev_alloc<int>(S);
  im.y_id = dev_alloc<int>(S);
  PKB_CUDA(cudaMemsetAsync(im.cache_tm, 0, S * L_ * kDModel * kTimeCtx * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_h, 0, S * kPredL * kPredH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_c, 0, S * kPredL * kPredH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_g, 0, S * kPredH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_proj, 0, S * kJointH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.n_emitted, 0, S * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.y_id, 0, S * 4, st_));

  im.Bcap = opt_.max_streams;
  im.Mcap = opt_.max_rows > 0 ? opt_.max_rows : std::max(64, 8 * opt_.max_streams);
  im.Mcap = std::max(im.Mcap, kMaxTq + 2);
  im.T3cap = im.Mcap + kDropPre * im.Bcap;
  im.Mcap = std::max(im.Mcap, kMaxTq);
  im.T2cap = 2 * im.T3cap + im.Bcap;
  const int rows_dec = std::max(im.Bcap, 64);
  im.a_sub1 = make_act(im.T2cap * 32, kSubCh, split, st_);
  im.a_sub2 = make_act(im.T3cap * 16, kSubCh, split, st_);
  im.a_sub3 = make_act(im.T3cap, 16 * kSubCh, split, st_);
  im.a_ln = make_act(std::max(im.Mcap, rows_dec), kDModel, split, st_);
  im.a_ff = make_act(im.Mcap, kFF, split, st_);
  // the decode side (joint / predictor) always runs on split hi+lo operands: it is a few percent of the step's FLOPs and
  // keeps the logits fp32-grade, so bf16 mode differs from the oracle only through the encoder output
  im.a_xf = make_act(std::max(im.Mcap, rows_dec), kDModel, true, st_);
  im.a_hid = make_act(rows_dec, kJointH, true, st_);
  im.a_pred = make_act(rows_dec, 2 * kPredH, true, st_);
  im.a_g = make_act(rows_dec, kPredH, true, st_);
  im.a_imp = make_act(kCacheS, kDModel, split, st_);
  im.a_pos = make_act(kPosRows, kDModel, true, st_);
  im.x = dev_alloc<float>((size_t)im.Mcap * kDModel);
  im.q = dev_alloc<float>((size_t)im.Mcap * kDModel);
  if (im.attn_mma) {
    const size_t q_rows = ((size_t)im.Mcap + 127) / 128 * 128;
    im.q_plane = (long long)q_rows * kDModel;
    im.q_bf16 = dev_alloc<__nv_bfloat16>((size_t)2 * im.q_plane);
    PKB_CUDA(cudaMemsetAsync(im.q_bf16, 0, (size_t)2 * im.q_plane * 2, st_));
    make_tensor_map_2d(&im.map_qv, im.q_bf16 + im.q_plane, q_rows, kDModel, kDModel, 128);
    im.g_pos = dev_alloc<__nv_bfloat16>(q_rows * kHeads * kPosRowsPad);
  }
  im.cglu = dev_alloc<float>((size_t)im.Mcap * kDModel);
  im.a_y1 = make_act(im.T2cap * 32, kSubCh, split, st_);
  im.enc_proj = dev_alloc<float>((size_t)std::max(im.Mcap, rows_dec) * kJointH);
  im.logits = dev_alloc<float>((size_t)rows_dec * kJointOut);
  im.gates = dev_alloc<float>((size_t)rows_dec * 4 * kPredH);
  im.part_rows_split = std::min(im.Mcap, 2048);      // split-K (up to 4 ways) is only used for small batched passes
  im.part_rows = std::max(2 * im.Mcap, 4 * im.part_rows_split) / 4;   // row stride between 4 splits; 1-2 splits may use Mcap-row slabs
  im.part_ws = dev_alloc<float>((size_t)4 * im.part_rows * kDModel);
  im.part_val = dev_alloc<float>((size_t)rows_dec * kArgmaxParts);
  im.part_idx = dev_alloc<int>((size_t)rows_dec * kArgmaxParts);
  im.dur_logits = dev_alloc<float>((size_t)rows_dec * kNDur);
  im.enc_out = dev_alloc<float>((size_t)im.Bcap * kDModel * kMaxTq);
  im.ppos_tmp = dev_alloc<float>((size_t)kPosRows * kDModel);
  im.scratch_f32_elems = (size_t)L_ * kCacheS * kDModel;     // one stream's contract cache (import / export staging)
  im.scratch_f32 = dev_alloc<float>(im.scratch_f32_elems);
  const size_t nints = (size_t)kNumBatchFields * im.Bcap + 3 * (im.Bcap + 1);
  im.batch_ints = dev_alloc<int>(nints);
  host_alloc(&im.batch_ints_host, nints * sizeof(int));
  im.row_entry = dev_alloc<int>(im.Mcap);
  im.row_pos = dev_alloc<int>(im.Mcap);
  im.rowmap3 = dev_alloc<int>(im.T3cap);
  im.imp_row_entry = dev_alloc<int>(kCacheS);
  im.imp_row_pos = dev_alloc<int>(kCacheS);
  fill_import_rows_kernel<<<1, kCacheS, 0, st_>>>(im.imp_row_entry, im.imp_row_pos, kCacheS);
  im.t_cur = dev_alloc<int>(im.Bcap); im.n_sym = dev_alloc<int>(im.Bcap); im.active = dev_alloc<int>(im.Bcap);
  im.emit_tok = dev_alloc<int>(im.Bcap); im.pred_rowmap = dev_alloc<int>(im.Bcap);
  im.n_steps = dev_alloc<int>((size_t)im.Bcap * (1 + kMaxStepsOffline * 3));
  im.steps = im.n_steps + im.Bcap;
  im.counters = dev_alloc<int>(4);
  im.force_toks = dev_alloc<int>(im.Bcap);
  host_alloc(&im.res_host, (size_t)im.Bcap * (1 + kMaxStepsOffline * 3) * sizeof(int));
  host_alloc(&im.counters_host, 4 * sizeof(int));
  for (auto& e : im.dec_events) PKB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  // audio: per-stream device buffers + one staging area for batched host pushes (8192 samples per stream per push)
  im.audio_buf = dev_alloc<float>((size_t)S * kAudioCap);
  im.audio_tmp = dev_alloc<float>((size_t)S * kAudioCap);
  im.audio_stage_cap = (size_t)im.Bcap * 8192;
  im.audio_stage = dev_alloc<float>(im.audio_stage_cap);
  host_alloc(&im.audio_stage_host, im.audio_stage_cap * sizeof(float));
  im.push_meta = dev_alloc<int>((size_t)3 * im.Bcap);
  host_alloc(&im.push_meta_host, (size_t)3 * im.Bcap * sizeof(int));
  im.segs_dev = dev_alloc<FrontSegment>(im.Bcap);
  host_alloc(&im.segs_host, (size_t)im.Bcap * sizeof(FrontSegment));
  im.fprefix_dev = dev_alloc<int>(im.Bcap + 1);
  host_alloc(&im.fprefix_host, (size_t)(im.Bcap + 1) * sizeof(int));
  im.fstage_cap = (size_t)kNMels * std::max(512, 64 * im.Bcap);      // one steady-state chunk (57 frames) per stream, >= 2 maximal pushes
  im.fstage_dev = dev_alloc<float>(im.fstage_cap);
  host_alloc(&im.fstage_host, im.fstage_cap * sizeof(float));
  im.fpush_cap = 2 * im.Bcap + 16;
  im.fpush_dev = dev_alloc<FeatPush>(im.fpush_cap);
  host_alloc(&im.fpush_host, (size_t)im.fpush_cap * sizeof(FeatPush));
  im.meta2 = dev_alloc<int>(2);
  im.guard_dev = dev_alloc<int>(4);
  host_alloc(&im.guard_host, 4 * sizeof(int));

  // ---- projected relative-position table per layer:  P_l[r] = linear_pos_l(pe[r]),  r in [-kPosNeg, kPosRows-kPosNeg)
  // pe[r][2i] = sin(r * div_i), pe[r][2i+1] = cos(r * div_i), div_i = exp(-(ln 1e4) * 2i / d_model)   (NeMo RelPositionalEncoding)
  {
    std::vector<float> pe((size_t)kPosRows * kDModel);
    for (int idx = 0; idx < kPosRows; ++idx) {
      const float r = (float)(idx - kPosNeg);
      for (int i = 0; i < kDModel / 2; ++i) {
        const float div = expf((float)(2 * i) * -(logf(10000.0f) / (float)kDModel));
        pe[(size_t)idx * kDModel + 2 * i] = sinf(r * div);
        pe[(size_t)idx * kDModel + 2 * i + 1] = cosf(r * div);
      }
    }
    float* pe_dev = dev_upload(pe);
    const long long n = (long long)kPosRows * kDModel;
    f32_to_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st_>>>(pe_dev, kDModel, 1, kPosRows, kDModel, im.a_pos.out());
    WeightsFile wf(opt_.model_dir + "/weights.bin");
    for (int l = 0; l < L_; ++l) {
      GemmW wp = upload_gemm_w(wf.bf16("encoder.layers." + std::to_string(l) + ".self_attn.linear_pos.weight"), kDModel, kDModel);
      GemmArgs g;
      g.A = im.a_pos.ptr; g.lda = kDModel; g.a_lo_off = im.a_pos.lo_off;
      g.W = wp.w; g.M = kPosRows; g.N = kDModel; g.K = kDModel;
      g.epi.mode = EPI_F32; g.epi.out_f32 = im.ppos_tmp; g.epi.ldo = kDModel;
      gemm_simt(g, st_);
      if (im.attn_mma) {
        im.layers[l].ppos_n = dev_alloc<__nv_bfloat16>((size_t)kHeads * kPosRowsPad * kDHead);
        PKB_CUDA(cudaMemsetAsync(im.layers[l].ppos_n, 0, (size_t)kHeads * kPosRowsPad * kDHead * 2, st_));
        ppos_natural_kernel<<<kPosRows, 256, 0, st_>>>(im.ppos_tmp, im.layers[l].ppos_n);
        make_tensor_map_2d(&im.layers[l].ppos_map, im.layers[l].ppos_n, (uint64_t)kHeads * kPosRowsPad, kDHead, kDHead, 128);
      } else {
        im.layers[l].ppos_t = dev_alloc_bytes((size_t)kPosRows * kDModel * kv_elem);
        ppos_transpose_kernel<<<kPosRows, 256, 0, st_>>>(im.ppos_tmp, im.layers[l].ppos_t, 1);
      }
      PKB_CUDA(cudaStreamSynchronize(st_));
      dev_free(wp.w);
    }
    dev_free(pe_dev);
  }
}

// ------------------------------------------------------------------------------------------------ streams
int Engine::open_stream() {
  for (int i = 0; i < (int)streams_.size(); ++i)
    if (!streams_[i]->open) {
      streams_[i]->open = true;
      reset_stream(i);
      return i;
    }
  throw std::runtime_error("no free stream slot (max_streams=" + std::to_string(opt_.max_streams) + ")");
}

void Engine::close_stream(int sid) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size(), "bad stream id");
  streams_[sid]->open = false;
  streams_[sid]->offline = false;
  streams_[sid]->run_norm = false;
  streams_[sid]->pending.clear();
  streams_[sid]->audio.clear();
}

// forget the chunks a stream has queued but not processed (error recovery of the one-chunk-per-push legacy session)
void Engine::drop_pending(int sid) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size(), "bad stream id");
  Stream& s = *streams_[sid];
  while (!s.pending.empty()) { s.frames_written = std::min(s.frames_written, s.pending.back().f0); s.pending.pop_back(); }
}

void Engine::reset_stream(int sid) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Stream& s = *streams_[sid];
  Impl& im = *im_;
  s.frames_written = 0; s.pending.clear(); s.audio.clear(); s.audio_mode = false;   // (s.offline is a property of the stream: kept)
  s.sched_chunk = 0; s.has_norm = false;
  s.dev_off = 0; s.dev_fill = 0;
  s.cache_len = 0; s.head = 0; s.chunks = 0; s.tokens.clear(); s.token_frames.clear(); s.enc_frames = 0; s.last = ChunkResult();
  s.lf_store_off = -1; s.lf_store_T = 0;
  const size_t slot = (size_t)s.slot;
  PKB_CUDA(cudaMemsetAsync(im.cache_tm + slot * L_ * kDModel * kTimeCtx, 0, (size_t)L_ * kDModel * kTimeCtx * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_h + slot * kPredL * kPredH, 0, kPredL * kPredH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_c + slot * kPredL * kPredH, 0, kPredL * kPredH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.pred_g + slot * kPredH, 0, kPredH * 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.n_emitted + slot, 0, 4, st_));
  PKB_CUDA(cudaMemsetAsync(im.run_stats + slot * kRunStats, 0, kRunStats * 4, st_));      // (s.run_norm, the mode, is kept)
  s.needs_prime = true;     // primed (batched with every other freshly reset stream) right before its first decode
}

void Engine::set_stream_offline(int sid, bool offline) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Stream& s = *streams_[sid];
  PKB_CHECK(s.frames_written == 0 && s.chunks == 0, "set_stream_offline: only on a freshly opened / reset stream");
  s.offline = offline;
}

// Next chunk of the cache-aware schedule (tools/verify_nemo/streaming_encoder_reference.py:522-550): chunk 0 = frames [0,41);
// chunk k = [start-9, start+48), start = 17 + 24 (k-1).
static inline void sched_chunk_range(long long k, long long* lo, long long* hi) {
  const long long start = k == 0 ? 0 : 17 + 24 * (k - 1);
  *lo = k == 0 ? 0 : start - 9;
  *hi = start + (k == 0 ? 41 : 48);
}

// True while step() still has work for the stream: an explicit chunk, audio that yields at least one more frame, or frames
// already in the feature ring that complete the next scheduled chunk (one frontend pass can produce several chunks' worth of
// frames, step() cuts one chunk per stream: callers drain with `while (has_pending) step()`).
bool Engine::has_pending(int sid) const {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size(), "bad stream id");
  const Stream& s = *streams_[sid];
  if (!s.open) return false;
  if (!s.pending.empty() || !s.audio.empty() || s.dev_fill >= 400) return true;
  if (s.audio_mode) {
    long long lo, hi;
    sched_chunk_range(s.sched_chunk, &lo, &hi);
    return s.frames_written >= hi;
  }
  return false;
}
const std::vector<int>& Engine::tokens(int sid) const { return streams_[sid]->tokens; }
const ChunkResult& Engine::last_chunk(int sid) const { return streams_[sid]->last; }
int Engine::prime_now(int sid) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Stream& s = *streams_[sid];
  if (s.needs_prime) { s.needs_prime = false; prime_streams({sid}); }
  if (!s.tokens.empty()) return s.tokens.back();
  return tok_lang_ >= 0 ? tok_lang_ : tok_start_ >= 0 ? tok_start_ : kBlank;
}
int Engine::last_encoder_output(int sid, float* out, int cap_T) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size(), "bad stream id");
  const Stream& s = *streams_[sid];
  PKB_CHECK(s.last_entry >= 0 && s.last_out_T > 0, "last_encoder_output: the stream has not been through a batched pass (or only a whole-utterance one)");
  const int T = std::min(s.last.encoded_len, cap_T);
  std::vector<float> tmp((size_t)kDModel * s.last_out_T);
  PKB_CUDA(cudaMemcpyAsync(tmp.data(), im_->enc_out + (size_t)s.last_entry * kDModel * s.last_out_T, tmp.size() * sizeof(float),
                           cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
  for (int c = 0; c < kDModel; ++c)
    for (int t = 0; t < T; ++t) out[(size_t)c * T + t] = tmp[(size_t)c * s.last_out_T + t];
  return s.last.encoded_len;
}
const std::vector<int>& Engine::token_frames(int sid) const { return streams_[sid]->token_frames; }
long long Engine::encoder_frames_done(int sid) const { return streams_[sid]->enc_frames; }
int Engine::stable_prefix(int sid, int revision_window_ms) const {
  const Stream& s = *streams_[sid];
  const long long edge_ms = s.enc_frames * 80 - std::max(revision_window_ms, 0);
  // token_frames is non-decreasing: count the tokens emitted at or before the window's left edge
  return (int)(std::upper_bound(s.token_frames.begin(), s.token_frames.end(), edge_ms,
                                [](long long ms, int frame) { return ms < (long long)frame * 80; }) - s.token_frames.begin());
}
int Engine::cache_len(int sid) const { return streams_[sid]->cache_len; }
long long Engine::chunks_done(int sid) const { return streams_[sid]->chunks; }

std::string Engine::detokenize(const std::vector<int>& ids) const { return vocab_.decode(ids); }

// ------------------------------------------------------------------------------------------------ NaN / Inf guard
// The reference samples the first 4096 elements of encoder_output and of the two cache outputs after every guarded encoder step
// (parakeet_trt.cpp:913-1013, call sites :2449-2481).  Same census here, on the device: stage 0 = the stream's encoder_output of the
// last pass, 1 = its cache_last_channel (the 4 newest rows of layer 0 in the contract ring), 2 = its cache_last_time (layer 0).
Engine::GuardResult Engine::nan_guard(int sid, int stage) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Impl& im = *im_;
  const Stream& s = *streams_[sid];
  GuardResult r;
  GuardRows rows{};
  const void* base = nullptr;
  int is_bf16 = 0;
  if (stage == 0) {
    if (s.last_entry < 0 || s.last_out_T <= 0) return r;
    base = im.enc_out + (size_t)s.last_entry * kDModel * s.last_out_T;
    r.count = (size_t)kDModel * s.last_out_T;
    rows.n_rows = 1; rows.row_len = (int)std::min<size_t>(r.count, 4096); rows.off[0] = 0;
  } else if (stage == 1) {
    if (!im.acache) return r;
    base = im.acache;      // layer 0
    is_bf16 = opt_.precision == 1 ? 0 : 1;
    r.count = (size_t)L_ * kCacheS * kDModel;
    rows.n_rows = 4; rows.row_len = kDModel;
    for (int j = 0; j < 4; ++j) rows.off[j] = ((long long)s.slot * kRingCap + (s.head + kCacheS - 4 + j) % kRingCap) * kDModel;
  } else {
    base = im.cache_tm + (size_t)s.slot * L_ * kDModel * kTimeCtx;
    r.count = (size_t)L_ * kDModel * kTimeCtx;
    rows.n_rows = 1; rows.row_len = (int)std::min<size_t>(r.count, 4096); rows.off[0] = 0;
  }
  r.sample_n = (size_t)rows.n_rows * rows.row_len;
  const int init[4] = {0, 0, 0x7fffffff, 0x7fffffff};
  PKB_CUDA(cudaStreamSynchronize(st_));
  memcpy(im.guard_host, init, sizeof(init));
  PKB_CUDA(cudaMemcpyAsync(im.guard_dev, im.guard_host, sizeof(init), cudaMemcpyHostToDevice, st_));
  guard_scan_kernel<<<4, 256, 0, st_>>>(base, is_bf16, rows, im.guard_dev);
  PKB_CUDA(cudaGetLastError());
  ++launches_;
  PKB_CUDA(cudaMemcpyAsync(im.guard_host, im.guard_dev, sizeof(init), cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
  r.nan_count = im.guard_host[0]; r.inf_count = im.guard_host[1];
  r.first_nan = r.nan_count ? im.guard_host[2] : -1;
  r.first_inf = r.inf_count ? im.guard_host[3] : -1;
  return r;
}

// ------------------------------------------------------------------------------------------------ input queues
void Engine::queue_features(int sid, const float* feats, int T) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  PKB_CHECK(T >= 1 && T <= 256, "queue_features: 1 <= T <= 256 frames per chunk");
  Stream& s = *streams_[sid];
  Impl& im = *im_;
  PKB_CHECK(!s.audio_mode, "stream is in audio mode");
  {      // reject a chunk the encoder cannot take HERE, so that a queued chunk can always be processed (see step())
    const int Tq = sub_len(sub_len(sub_len(T))) - (s.offline ? 0 : kDropPre);
    if (s.offline) PKB_CHECK(Tq >= 1 && Tq <= kMaxTq, "offline push must hold 1..256 feature frames (got " + std::to_string(T) + ")");
    else PKB_CHECK(Tq >= kCacheDrop && Tq <= kMaxTq, "chunk must hold 33..256 feature frames (got " + std::to_string(T) + ")");
  }
  // all frame indices are absolute here; the ring index is taken (mod kFeatRing) only when the batch is uploaded
  PKB_CHECK(s.pending.empty() || s.frames_written + T - s.pending.front().f0 <= kFeatRing, "feature ring full: call step() first");
  const size_t need = (size_t)kNMels * T;
  if (im.fstage_inflight) { PKB_CUDA(cudaStreamSynchronize(st_)); im.fstage_inflight = false; im.fstage_used = 0; im.fpush_n = 0; im.fpush_max_T = 0; }
  if (im.fstage_used + need > im.fstage_cap || im.fpush_n == im.fpush_cap) flush_feature_stage(true);
  memcpy(im.fstage_host + im.fstage_used, feats, need * sizeof(float));
  FeatPush& d = im.fpush_host[im.fpush_n++];
  d.src_off = (long long)im.fstage_used;
  d.ring_off = (long long)s.slot * kFeatRing * kNMels;
  d.T = T;
  d.frame0 = (int)(s.frames_written % kFeatRing);
  im.fstage_used += need;
  im.fpush_max_T = std::max(im.fpush_max_T, T);
  s.pending.push_back(Entry{sid, s.frames_written, T});
  s.frames_written += T;
}

// Staged feature pushes -> feature rings: one H2D copy of the staging area, one copy of the descriptors, one transpose kernel.
void Engine::flush_feature_stage(bool sync) {
  Impl& im = *im_;
  if (im.fpush_n > 0 && !im.fstage_inflight) {
    PKB_CUDA(cudaMemcpyAsync(im.fstage_dev, im.fstage_host, im.fstage_used * sizeof(float), cudaMemcpyHostToDevice, st_));
    PKB_CUDA(cudaMemcpyAsync(im.fpush_dev, im.fpush_host, (size_t)im.fpush_n * sizeof(FeatPush), cudaMemcpyHostToDevice, st_));
    im.frontend.bins_to_frames_batch(im.fstage_dev, im.fpush_dev, im.fpush_n, im.fpush_max_T, im.feat_ring, kFeatRing, st_);
    ++launches_;
    im.fstage_inflight = true;
  }
  if (sync && im.fstage_inflight) {
    PKB_CUDA(cudaStreamSynchronize(st_));
    im.fstage_inflight = false; im.fstage_used = 0; im.fpush_n = 0; im.fpush_max_T = 0;
  }
}

void Engine::queue_audio(int sid, const float* pcm, size_t n) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Stream& s = *streams_[sid];
  PKB_CHECK(s.audio_mode || s.frames_written == 0, "stream already received features; reset it before pushing audio");
  s.audio_mode = true;
  s.audio.insert(s.audio.end(), pcm, pcm + n);     // moved to the device buffer by step() as room allows
}

// Batched push: `count` samples for each of n streams, source row i at src + i*stride (host or device memory).
// One H2D copy (host source) + one append kernel, no per-stream host work beyond bookkeeping.
void Engine::push_audio_batch(int n, const int* sids, const float* src, long long stride, int count, bool src_on_device,
                              bool internal) {
  Impl& im = *im_;
  PKB_CHECK(n >= 0 && n <= im.Bcap && count >= 0 && count <= 8192, "push_audio_batch: n <= max_streams, count <= 8192");
  if (n == 0 || count == 0) return;
  PKB_CUDA(cudaStreamSynchronize(st_));     // push_meta_host / staging reuse
  bool need_compact = false;
  for (int i = 0; i < n; ++i) {
    PKB_CHECK(sids[i] >= 0 && sids[i] < (int)streams_.size() && streams_[sids[i]]->open, "bad stream id");
    Stream& s = *streams_[sids[i]];
    PKB_CHECK(s.audio_mode || s.frames_written == 0, "stream already received features; reset it before pushing audio");
    PKB_CHECK(internal || s.audio.empty(), "push_audio_batch: stream has host-queued audio pending; call step() first");
    PKB_CHECK(s.dev_fill + count <= kAudioCap, "device audio buffer full: call step() before pushing more");
    if (s.dev_off + s.dev_fill + count > kAudioCap) need_compact = true;
  }
  if (need_compact) compact_audio();
  const float* dsrc = src;
  long long dstride = stride;
  if (!src_on_device) {
    PKB_CHECK((size_t)n * count <= im.audio_stage_cap, "push_audio_batch: staging too small");
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (stride == count) {
      const float* h = src;
      if (!pinned) { memcpy(im.audio_stage_host, src, (size_t)n * count * 4); h = im.audio_stage_host; }
      PKB_CUDA(cudaMemcpyAsync(im.audio_stage, h, (size_t)n * count * 4, cudaMemcpyHostToDevice, st_));
    } else {
      for (int i = 0; i < n; ++i) memcpy(im.audio_stage_host + (size_t)i * count, src + (size_t)i * stride, (size_t)count * 4);
      PKB_CUDA(cudaMemcpyAsync(im.audio_stage, im.audio_stage_host, (size_t)n * count * 4, cudaMemcpyHostToDevice, st_));
    }
    dsrc = im.audio_stage;
    dstride = count;
  }
  int* h = im.push_meta_host;
  for (int i = 0; i < n; ++i) {
    Stream& s = *streams_[sids[i]];
    h[i] = s.slot;
    h[im.Bcap + i] = s.dev_off + s.dev_fill;
    h[2 * im.Bcap + i] = count;
    s.dev_fill += count;
    s.audio_mode = true;
  }
  PKB_CUDA(cudaMemcpyAsync(im.push_meta, h, (size_t)3 * im.Bcap * sizeof(int), cudaMemcpyHostToDevice, st_));
  launch_k(audio_append_kernel, dim3((count + 1023) / 1024, n), dim3(256), 0, st_, dsrc, dstride, im.audio_buf, (const int*)im.push_meta,
           (const int*)(im.push_meta + im.Bcap), (const int*)(im.push_meta + 2 * im.Bcap));
  ++launches_;
}

// move every stream's valid samples to the front of its device buffer (through a scratch copy: ranges may overlap)
void Engine::compact_audio() {
  Impl& im = *im_;
  PKB_CUDA(cudaStreamSynchronize(st_));
  int* h = im.push_meta_host;
  int n = 0;
  for (auto& sp : streams_) {
    Stream& s = *sp;
    if (!s.open || s.dev_off == 0) continue;
    h[n] = s.slot; h[im.Bcap + n] = s.dev_off; h[2 * im.Bcap + n] = s.dev_fill;
    s.dev_off = 0;
    ++n;
  }
  if (n == 0) return;
  PKB_CUDA(cudaMemcpyAsync(im.push_meta, h, (size_t)3 * im.Bcap * sizeof(int), cudaMemcpyHostToDevice, st_));
  audio_move_kernel<<<dim3(kAudioCap / 1024, n), 256, 0, st_>>>(im.audio_buf, im.audio_tmp, im.push_meta, im.push_meta + im.Bcap,
                                                                im.push_meta + 2 * im.Bcap, 0);
  audio_move_kernel<<<dim3(kAudioCap / 1024, n), 256, 0, st_>>>(im.audio_buf, im.audio_tmp, im.push_meta, im.push_meta + im.Bcap,
                                                                im.push_meta + 2 * im.Bcap, 1);
  PKB_CUDA(cudaGetLastError());
  launches_ += 2;
  PKB_CUDA(cudaStreamSynchronize(st_));
}

void Engine::set_feature_norm(int sid, const float* mean128, const float* std128) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Stream& s = *streams_[sid];
  if (!mean128 || !std128) { s.has_norm = false; return; }
  float* dst = im_->norm_stats + (size_t)s.slot * 2 * kNMels;
  PKB_CUDA(cudaMemcpy(dst, mean128, kNMels * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(dst + kNMels, std128, kNMels * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaDeviceSynchronize());   // pageable H2D: staged != landed (see dev_upload)
  s.has_norm = true;
}

void Engine::set_feature_norm_running(int sid, bool on) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Stream& s = *streams_[sid];
  PKB_CHECK(s.frames_written == 0, "set_feature_norm_running: only on a freshly opened / reset stream");
  s.run_norm = on;
  PKB_CUDA(cudaMemsetAsync(im_->run_stats + (size_t)s.slot * kRunStats, 0, kRunStats * 4, st_));
}

// audio -> feature rings for every stream in audio mode (one launch); chunks are then cut by the schedule in step()
void Engine::frontend_pass() {
  Impl& im = *im_;
  if (im.frontend_inflight) { PKB_CUDA(cudaStreamSynchronize(st_)); im.frontend_inflight = false; }      // (only after a failed pass)
  // 1. top up device buffers from host overflow FIFOs (queue_audio path)
  for (int sid = 0; sid < (int)streams_.size(); ++sid) {
    Stream& s = *streams_[sid];
    if (!s.open || s.audio.empty()) continue;
    if (s.dev_off + s.dev_fill + 4096 > kAudioCap && s.dev_off > 0) compact_audio();
    const int room = kAudioCap - (s.dev_off + s.dev_fill);
    const int n = (int)std::min<size_t>(std::min<size_t>(s.audio.size(), (size_t)room), 8192);
    if (n <= 0) continue;
    std::vector<float> part(s.audio.begin(), s.audio.begin() + n);
    s.audio.erase(s.audio.begin(), s.audio.begin() + n);
    const int one = sid;
    push_audio_batch(1, &one, part.data(), n, n, false, true);
  }
  // 2. one log-mel launch over every stream that holds at least one complete frame and has ring room
  int n_segs = 0, total_frames = 0;
  bool any_running = false;
  std::vector<std::pair<int, int>> done;   // (sid, frames)
  for (int sid = 0; sid < (int)streams_.size(); ++sid) {
    Stream& s = *streams_[sid];
    if (!s.open || !s.audio_mode || s.dev_fill < 400) continue;
    int frames = (s.dev_fill - 400) / 160 + 1;
    // frames not yet consumed by the schedule must stay resident in the feature ring
    const long long next_start = s.sched_chunk == 0 ? 0 : 17 + 24 * (s.sched_chunk - 1);
    const long long keep_from = std::max(0LL, next_start - 9);
    const long long room = kFeatRing - (s.frames_written - keep_from);
    if (room <= 0) continue;
    frames = (int)std::min<long long>(frames, room);
    FrontSegment& sg = im.segs_host[n_segs];
    sg.audio_off = (long long)s.slot * kAudioCap + s.dev_off;
    sg.out_off = (long long)s.slot * kFeatRing * kNMels;
    sg.out_stride = kNMels;
    sg.ring_cap = kFeatRing;
    sg.frame0 = (int)(s.frames_written % kFeatRing);
    sg.norm_off = (s.has_norm && !s.run_norm) ? s.slot * 2 * kNMels : -1;
    sg.run_state = s.run_norm ? 1 + s.slot * kRunStats : 0;
    any_running = any_running || s.run_norm;
    im.fprefix_host[n_segs] = total_frames;
    total_frames += frames;
    ++n_segs;
    done.emplace_back(sid, frames);
  }
  if (n_segs == 0) return;
  im.fprefix_host[n_segs] = total_frames;
  PKB_CUDA(cudaMemcpyAsync(im.segs_dev, im.segs_host, n_segs * sizeof(FrontSegment), cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaMemcpyAsync(im.fprefix_dev, im.fprefix_host, (n_segs + 1) * sizeof(int), cudaMemcpyHostToDevice, st_));
  const int pi = prof_begin(2, (double)total_frames * (kNMels * 4.0 + 160 * 4.0));      // 640 B of new samples read + 512 B written per frame
  im.frontend.logmel(im.audio_buf, im.segs_dev, im.fprefix_dev, n_segs, total_frames, im.feat_ring, im.norm_stats, sm_count_, st_);
  ++launches_;
  prof_end(pi);
  if (any_running) { im.frontend.running_norm(im.feat_ring, im.segs_dev, im.fprefix_dev, n_segs, im.run_stats, st_); ++launches_; }
  for (auto& d : done) {
    Stream& s = *streams_[d.first];
    s.dev_off += d.second * 160;     // stays even
    s.dev_fill -= d.second * 160;
    s.frames_written += d.second;
  }
  // segs_host / fprefix_host are rewritten only by the next frontend_pass; the batched pass that normally follows ends with a stream
  // synchronisation, and step() synchronises itself when no chunk was ready (no host round trip in the middle of a step)
  im.frontend_inflight = true;
}

// ------------------------------------------------------------------------------------------------ batching
BatchDev Engine::upload_batch(const std::vector<Entry>& entries) {
  Impl& im = *im_;
  const int B = (int)entries.size();
  PKB_CHECK(B <= im.Bcap, "batch larger than max_streams");
  int* h = im.batch_ints_host;
  const int C = im.Bcap;
  int *slot = h, *T = h + C, *f0 = h + 2 * C, *T1 = h + 3 * C, *T2 = h + 4 * C, *T3 = h + 5 * C, *Tq = h + 6 * C, *qlen = h + 7 * C,
      *len = h + 8 * C, *head = h + 9 * C, *tenc = h + 10 * C, *drop = h + 11 * C, *offl = h + 12 * C;
  int* off2 = h + kNumBatchFields * C;
  int* off3 = off2 + (C + 1);
  int* roff = off3 + (C + 1);
  BatchDev b;
  b.B = B;
  off2[0] = off3[0] = roff[0] = 0;
  for (int i = 0; i < B; ++i) {
    const Stream& s = *streams_[entries[i].sid];
    slot[i] = s.slot;
    T[i] = entries[i].T;
    f0[i] = (int)(entries[i].f0 % kFeatRing);
    T1[i] = sub_len(T[i]);
    T2[i] = sub_len(T1[i]);
    T3[i] = sub_len(T2[i]);
    // streaming (forward_for_export): drop_extra_pre_encoded tokens go, cache_drop_size more are recomputed next chunk,
    // the first valid_out_len rows are emitted.  offline (the reference's `encoder` engine, contract.json:67-96): every
    // token of the push is kept and emitted, no cache.
    offl[i] = s.offline ? 1 : 0;
    drop[i] = s.offline ? 0 : kDropPre;
    Tq[i] = T3[i] - drop[i];
    if (s.offline)
      PKB_CHECK(Tq[i] >= 1 && Tq[i] <= kMaxTq, "offline push must hold 1..256 feature frames (got " + std::to_string(T[i]) + ")");
    else
      PKB_CHECK(Tq[i] >= kCacheDrop && Tq[i] <= kMaxTq, "chunk must hold 33..256 feature frames (got " + std::to_string(T[i]) + ")");
    qlen[i] = Tq[i];
    len[i] = s.offline ? 0 : s.cache_len;
    head[i] = s.head;
    tenc[i] = s.offline ? Tq[i] : std::min(qlen[i], kValidOut);
    b.max_tenc = std::max(b.max_tenc, tenc[i]);
    off2[i + 1] = off2[i] + T2[i];
    off3[i + 1] = off3[i] + T3[i];
    roff[i + 1] = roff[i] + Tq[i];
    b.max_Tq = std::max(b.max_Tq, Tq[i]);
  }
  b.M = roff[B]; b.sumT2 = off2[B]; b.sumT3 = off3[B];
  PKB_CHECK(b.M <= im.Mcap && b.sumT3 <= im.T3cap && b.sumT2 <= im.T2cap, "batch exceeds row capacity");
  const size_t nints = (size_t)kNumBatchFields * C + 3 * (C + 1);
  PKB_CUDA(cudaMemcpyAsync(im.batch_ints, h, nints * sizeof(int), cudaMemcpyHostToDevice, st_));
  int* d = im.batch_ints;
  b.slot = d; b.T = d + C; b.f0 = d + 2 * C; b.T1 = d + 3 * C; b.T2 = d + 4 * C; b.T3 = d + 5 * C; b.Tq = d + 6 * C;
  b.qlen = d + 7 * C; b.len = d + 8 * C; b.head = d + 9 * C; b.drop = d + 11 * C; b.offline = d + 12 * C;
  b.off2 = d + kNumBatchFields * C; b.off3 = b.off2 + (C + 1); b.row_off = b.off3 + (C + 1);
  b.row_entry = im.row_entry; b.row_pos = im.row_pos; b.rowmap3 = im.rowmap3;
  return b;
}

// GEMM on an activation buffer and a weight, choosing the backend
static int g_tc_site = 0;        // bring-up aid: PARAKEET_B200_TC_MASK disables the tensor-core path per call-site group
static int tc_mask() {
  static int m = -2;
  if (m == -2) { const char* v = getenv("PARAKEET_B200_TC_MASK"); m = v ? atoi(v) : -1; }
  return m;
}
static void run_gemm(Engine* eng, const EngineOptions& opt, cudaStream_t st, long long* launches, const ActBuf& a, int lda_override,
                     const GemmW& w, int M, const int* M_dev, const EpiParams& epi) {
  (void)eng;
  GemmArgs g;
  g.A = a.ptr;
  g.lda = lda_override > 0 ? lda_override : a.K;
  g.a_lo_off = a.lo_off;
  g.W = w.w;
  g.M = M; g.N = w.N; g.K = w.K;
  g.M_dev = M_dev;
  g.map_w32 = &w.map32;
  g.epi = epi;
  ++*launches;
  bool want_tc = opt.gemm_backend == 2 || (opt.gemm_backend == 0 && M > 16);
  if (tc_mask() >= 0 && !(tc_mask() & g_tc_site)) want_tc = false;
  if (want_tc && lda_override <= 0 && gemm_tc_supported(g)) {
    const int pi = eng->prof_begin(0, 2.0 * (double)M * w.N * w.K);   // ALGORITHMIC flops (the split pass is not counted twice)
    gemm_tc(g, a.map, w.map, st);
    eng->prof_end(pi);
  } else {
    gemm_simt(g, st);
  }
}
#define RUN_GEMM(a, w, M, Mdev, epi) run_gemm(this, opt_, st_, &launches_, a, 0, w, M, Mdev, epi)

void Engine::run_encoder(const BatchDev& b, const LongForm* lf) {
  Impl& im = *im_;
  const bool split = opt_.precision == 1;
  launch_build_rows(b, st_); ++launches_;
  // ---- pre-encode ----
  launch_subsample_stage1(b, lf ? im.lf_feat : im.feat_ring, lf ? (int)im.lf_feat_frames : kFeatRing, im.sub, im.a_sub1.out(), st_); ++launches_;
  g_tc_site = 1;
  { EpiParams e; e.mode = EPI_BIAS_RELU_ACT; e.out_act = im.a_y1.ptr; e.lda_out = kSubCh; e.lo_off_out = im.a_y1.lo_off; e.bias = im.sub_pw1_b;
    RUN_GEMM(im.a_sub1, im.sub_pw1, b.sumT2 * 32, nullptr, e); }
  launch_subsample_stage2(b, im.a_y1.out(), im.sub, im.a_sub2.out(), st_); ++launches_;
  g_tc_site = 2;
  { EpiParams e; e.mode = EPI_BIAS_RELU_ACT; e.out_act = im.a_sub3.ptr; e.lda_out = kSubCh; e.lo_off_out = im.a_sub3.lo_off;
    e.bias = im.sub_pw2_b;
    RUN_GEMM(im.a_sub2, im.sub_pw2, b.sumT3 * 16, nullptr, e); }
  g_tc_site = 4;
  { EpiParams e; e.mode = EPI_BIAS_ROWMAP_F32; e.out_f32 = im.x; e.ldo = kDModel; e.bias = im.sub_out_b; e.row_map = b.rowmap3;
    RUN_GEMM(im.a_sub3, im.sub_out, b.sumT3, nullptr, e); }
  // ---- conformer layers ----
  g_tc_site = 8;
  const int M = b.M;
  // Residual GEMMs (N = 1024).  Large batches add into x in the epilogue.  Small batches leave too few 128 x 128 tiles for
  // 148 SMs, so the k-range is split over more CTAs; the partial sums go to a workspace and the LayerNorm that follows adds
  // them to x in a fixed order (deterministic, no atomics).
  auto residual_gemm = [&](const ActBuf& act, const GemmW& wt, float scale) -> LnResidual {
    const bool tc = (opt_.gemm_backend == 2 || (opt_.gemm_backend == 0 && M > 16)) && tc_mask() < 0;
    const int tiles = ((M + 127) / 128) * (wt.N / 128);
    int splits = tiles > 0 ? std::min(4, sm_count_ / tiles) : 1;
    splits = std::min(splits, wt.K / 64 / 2);
    static const bool allow = [] { const char* v = getenv("PARAKEET_B200_SPLITK"); return !(v && v[0] == '0'); }();
    // (splits == 1 still defers the add: the GEMM epilogue becomes write-only instead of a latency-bound read-modify-write of
    // x, measured 4 % faster per step at 1024 streams)
    static const bool defer_all = [] { const char* v = getenv("PARAKEET_B200_DEFER"); return !(v && v[0] == '0'); }();
    splits = std::max(splits, 1);
    if (splits > 1 && M > im.part_rows_split) splits = 1;
    // Small batches are bound by L2 -> SM operand traffic (a 128 x 128 tile moves 32 KB per k-block for 128 x 128 x 64 MACs; at 128
    // streams a K = 4096 residual GEMM pulls 97 MB through L2 in ~10 us): 128 x 256 tiles move 48 KB for twice the MACs.  Taken when
    // the wider tiles, split up to 4 ways, still give most SMs a unit (PARAKEET_B200_PART_WIDE=0: always 128-wide).
    static const bool wide_allowed = [] { const char* v = getenv("PARAKEET_B200_PART_WIDE"); return !(v && v[0] == '0'); }();
    int part_wide = 0;
    if (tc && allow && wide_allowed && wt.N % 256 == 0 && M <= im.part_rows_split) {
      const int tiles_w = ((M + 127) / 128) * (wt.N / 256);
      int splits_w = std::max(1, std::min(4, sm_count_ / std::max(tiles_w, 1)));
      splits_w = std::max(1, std::min(splits_w, wt.K / 64 / 2));
      // (measured, gpurun r3a: 256 streams 5.87 -> 5.37 ms, 192 streams 4.69 -> 4.62; at 128 streams the wide tiling leaves 96 units
      // against 144 narrow ones and loses 5 %: wide only when it does not shrink the unit count)
      if (tiles_w * splits_w * 5 >= sm_count_ * 3 && tiles_w * splits_w >= tiles * splits) { part_wide = 1; splits = splits_w; }
    }
    int stride_rows = im.part_rows, pair_split = 0;
    // Small batch, long K (the FFN-down projections): 256 x 256 pair tiles split up to 4 ways halve the operand bytes per MAC once
    // more (M = 768: 12 tiles x 4 splits = 48 pairs, 49 MB through L2 instead of 97 MB) -- but leave a third of the SMs without a
    // unit and make every unit's k-loop longer: measured slower at 128 ... 320 streams (3.99 vs 3.93, 4.67 vs 4.55, 5.60 vs 5.34,
    // 6.03 vs 5.97 ms; gpurun r3b), so OFF unless PARAKEET_B200_PAIR_SMALL=1.
    static const bool pair_small_allowed = [] { const char* v = getenv("PARAKEET_B200_PAIR_SMALL"); return v && v[0] == '1'; }();
    if (tc && allow && pair_small_allowed && wt.N % 256 == 0 && wt.K >= 2048 && M >= 256 && M <= im.part_rows_split) {
      const int tiles_p = ((M + 255) / 256) * (wt.N / 256), pairs = sm_count_ / 2;
      int splits_p = std::max(1, std::min(4, pairs / std::max(tiles_p, 1)));
      splits_p = std::max(1, std::min(splits_p, wt.K / 64 / 2));
      if (splits_p >= 2 && tiles_p * splits_p * 5 >= pairs * 3) { splits = splits_p; pair_split = 2; part_wide = 0; }
    }
    if (tc && allow && splits == 1) {      // large batch on the CTA-pair kernel: a 2-way k-split can fill its last wave
      const int ps = gemm_tc_pair_splits(M, wt.N, wt.K);
      if (ps == 2 && 2 * (size_t)im.Mcap <= (size_t)4 * im.part_rows) { splits = 2; stride_rows = im.Mcap; pair_split = 1; }
    }
    if (tc && allow && (splits >= 2 || defer_all) && wt.N == kDModel) {
      EpiParams e; e.mode = EPI_PARTIAL_F32; e.out_f32 = im.part_ws; e.ldo = kDModel; e.splits = splits; e.part_rows = stride_rows;
      e.pair_split = pair_split;
      e.part_wide = (part_wide && !pair_split) ? 1 : 0;
      // bf16 mode keeps the deferred branch outputs in bf16 (they are O(|x|/3) and join a stream whose GEMM operands are
      // rounded to bf16 anyway; parity set unchanged at 256/256, 2 % less time per step); precise mode keeps f32
      static const bool pb = [] { const char* v = getenv("PARAKEET_B200_PART_BF16"); return !(v && v[0] == '0'); }();
      e.part_bf16 = (pb && !split) ? 1 : 0;
      RUN_GEMM(act, wt, M, nullptr, e);
      return LnResidual{im.part_ws, splits, (long long)stride_rows * kDModel, scale, e.part_bf16};
    }
    EpiParams e; e.mode = EPI_RESADD_F32; e.out_f32 = im.x; e.ldo = kDModel; e.scale = scale;
    RUN_GEMM(act, wt, M, nullptr, e);
    return LnResidual{nullptr, 0, 0, 0.f, 0};
  };
  // q | k | v projection + attention of one layer over [per-stream K/V ring (256 cached rows) || the chunk's rows]
  // One-stream latency path (M <= 16, bf16 mode, CUDA-core GEMMs): the LayerNorms that feed exactly one projection are computed
  // inside that projection's kernel (gemm_simt_ln) instead of as launches of their own (PARAKEET_B200_LN_FUSE=0 restores them).
  static const bool ln_fuse_allowed = [] { const char* v = getenv("PARAKEET_B200_LN_FUSE"); return !(v && v[0] == '0'); }();
  const bool fuse_ln = ln_fuse_allowed && !split && !lf && M <= 16 && opt_.gemm_backend != 2;
  auto gemm_ln = [&](const float* gamma, const float* beta, const AcacheOut* ac, const GemmW& wt, const EpiParams& e) {
    GemmArgs g;
    g.A = im.a_ln.ptr; g.lda = kDModel; g.W = wt.w; g.M = M; g.N = wt.N; g.K = wt.K; g.epi = e;
    LnFuse f{im.x, gamma, beta, ac ? *ac : AcacheOut{}, ac ? 1 : 0};
    ++launches_;
    gemm_simt_ln(g, f, st_);
  };
  auto attention_streaming = [&](int l, const LayerW& w, char* kr, char* vr, const AcacheOut* fused_ac, bool fused) {
    { EpiParams e; e.mode = EPI_QKV; e.out_f32 = im.q; e.ldo = kDModel; e.row_entry = b.row_entry; e.row_pos = b.row_pos;
      e.entry_slot = b.slot; e.entry_head = b.head; e.kring = kr; e.vring = vr; e.kv_f32 = split ? 1 : 0;
      e.k_natural = im.attn_mma ? 1 : 0;
      if (im.attn_mma) { e.q_bf16 = im.q_bf16; e.q_plane = im.q_plane; e.bias_u = w.bias_u; e.bias_v = w.bias_v; }
      if (fused) gemm_ln(w.n_att_g, w.n_att_b, fused_ac, w.qkv, e);
      else RUN_GEMM(im.a_ln, w.qkv, M, nullptr, e); }
    if (im.attn_mma) {
      // position scores for every (row, head): G[m][h][r] = (q_m + pos_bias_v)[h] . P_h[r] -- 8 batched [M,128] x [320,128]^T
      // problems in one tcgen05 launch (the table is the same for every stream, so this does not belong in the per-stream kernel)
      { GemmArgs g;
        g.A = im.q_bf16 + im.q_plane; g.lda = kDModel; g.W = w.ppos_n; g.M = M; g.N = kPosRowsPad; g.K = kDHead;
        g.batch = kHeads; g.a_col_stride = kDHead; g.w_row_stride = kPosRowsPad; g.out_col_stride = kPosRowsPad;
        g.epi.mode = EPI_ACT; g.epi.out_act = im.g_pos; g.epi.lda_out = kHeads * kPosRowsPad;
        gemm_tc(g, im.map_qv, w.ppos_map, st_); ++launches_; }
      AttnMmaArgs a; a.q_bf16 = im.q_bf16; a.q_plane = im.q_plane; a.g_pos = im.g_pos; a.ctx = im.a_ln.out();
      a.map_k = &im.map_k; a.map_v = &im.map_v; a.layer = l; a.n_slots = opt_.max_streams;
      a.map_k32 = &im.map_k32; a.map_v32 = &im.map_v32; a.map_k8 = &im.map_k8; a.map_v8 = &im.map_v8;
      // algorithmic bytes: the valid K and V rows of every (stream, head): 2 x (256 + Tq) x 128 x 2 B (steady state; early chunks less)
      const int pi = prof_begin(1, (double)b.B * kHeads * 2.0 * (kCacheS + b.max_Tq) * kDHead * 2.0);
      launch_attention_mma(b, a, st_); ++launches_;
      prof_end(pi);
    } else {
      AttnArgs a; a.q = im.q; a.kring = kr; a.vring = vr; a.ppos_t = w.ppos_t; a.kv_f32 = 1; a.bias_u = w.bias_u;
      a.bias_v = w.bias_v; a.ctx = im.a_ln.out();
      launch_attention(b, a, st_); ++launches_;
    }
  };
  // whole-utterance attention of one layer: q | k | v rows stay in one [M,3072] buffer (no rings), the layer's projected position
  // table linear_pos(pe) over 2Tm-1 relative positions is one more GEMM, the score tiles are formed in lf_attention
  auto attention_whole_utterance = [&](int l, const LayerW& w, const LongForm& lf) {
    { EpiParams e;
      if (split) { e.mode = EPI_F32; e.out_f32 = (float*)im.lf_qkv; e.ldo = 3 * kDModel; }
      else { e.mode = EPI_ACT; e.out_act = (__nv_bfloat16*)im.lf_qkv; e.lda_out = 3 * kDModel; }
      RUN_GEMM(im.a_ln, w.qkv, M, nullptr, e); }
    { EpiParams e;
      if (split) { e.mode = EPI_F32; e.out_f32 = (float*)im.lf_ppos; e.ldo = kDModel; }
      else { e.mode = EPI_ACT; e.out_act = (__nv_bfloat16*)im.lf_ppos; e.lda_out = kDModel; }
      RUN_GEMM(im.lf_pos_act, im.lf_wpos[l], 2 * lf.Tm - 1, nullptr, e); }
    LfAttnArgs a;
    if (split) { a.qkv_f32 = (const float*)im.lf_qkv; a.ppos_f32 = (const float*)im.lf_ppos; }
    else { a.qkv_bf16 = (const __nv_bfloat16*)im.lf_qkv; a.ppos_bf16 = (const __nv_bfloat16*)im.lf_ppos; }
    a.Tm = lf.Tm; a.max_T = b.max_Tq; a.bias_u = w.bias_u; a.bias_v = w.bias_v; a.ctx = im.a_ln.out();
    if (!split) {
      if (l == 0) {      // same buffers for every layer: the maps only depend on this call's row counts
        make_tensor_map_2d(&im.lf_map_qkv, im.lf_qkv, (uint64_t)M, 3 * kDModel, 3 * kDModel, 64);
        make_tensor_map_2d(&im.lf_map_pos, im.lf_ppos, (uint64_t)(2 * lf.Tm - 1), kDModel, kDModel, 128);
        if (im.lf_attn_kind == 2) {
          make_tensor_map_2d(&im.lf_map_q, im.lf_qplanes, (uint64_t)2 * im.lf_q_rows, kDModel, kDModel, 128);
          make_tensor_map_2d(&im.lf_map_pos64, im.lf_ppos, (uint64_t)(2 * lf.Tm - 1), kDModel, kDModel, 64);
          make_tensor_map_2d(&im.lf_map_vt, im.lf_vt, (uint64_t)kDModel, (uint64_t)im.lf_ldv, (uint64_t)im.lf_ldv, 128);
        }
      }
      a.map_qkv = im.lf_attn_kind >= 1 ? &im.lf_map_qkv : nullptr; a.map_pos = im.lf_attn_kind >= 1 ? &im.lf_map_pos : nullptr;
    }
    double pairs = 0.0;      // sum over utterances of T^2 (host copy of the batch fields)
    for (int i = 0; i < b.B; ++i) { const double t = im.batch_ints_host[6 * im.Bcap + i]; pairs += t * t; }
    // algorithmic FLOPs: content score, position score and value product, 128 MACs each per (query, key, head)
    if (!split && im.lf_attn_kind == 2) {
      // tcgen05 kernel: biased query planes + V^T once per layer, then S / position blocks / PV on the 5th-generation tensor cores
      LfTcArgs t;
      t.qkv = (const __nv_bfloat16*)im.lf_qkv; t.bias_u = w.bias_u; t.bias_v = w.bias_v;
      t.q_planes = im.lf_qplanes; t.q_plane_rows = im.lf_q_rows; t.q_plane = (long long)im.lf_q_rows * kDModel;
      t.vt = im.lf_vt; t.ldv = im.lf_ldv; t.Tm = lf.Tm; t.ctx = im.a_ln.ptr; t.ldc = kDModel;
      t.map_q = &im.lf_map_q; t.map_k = &im.lf_map_qkv; t.map_pos = &im.lf_map_pos64; t.map_vt = &im.lf_map_vt;
      launch_lf_prep(b, t, st_); ++launches_;
      const int pi = prof_begin(4, 2.0 * 3.0 * kDHead * kHeads * pairs);
      launch_lf_attention_tc(b, t, b.max_Tq, st_); ++launches_;
      prof_end(pi);
      return;
    }
    const int pi = prof_begin(4, 2.0 * 3.0 * kDHead * kHeads * pairs);
    launch_lf_attention(b, a, st_); ++launches_;
    prof_end(pi);
  };
  LnResidual res{};
  launch_layernorm(im.x, M, im.layers[0].n_ff1_g, im.layers[0].n_ff1_b, nullptr, nullptr, 0, im.a_ln.out(), nullptr, st_); ++launches_;
  const size_t kv_elem = split ? 4 : 2;
  for (int l = 0; l < L_; ++l) {
    const LayerW& w = im.layers[l];
    char* kr = (char*)im.kring + (size_t)l * im.ring_layer_elems * kv_elem;
    char* vr = (char*)im.vring + (size_t)l * im.ring_layer_elems * kv_elem;
    // FFN 1 (half-step residual)
    g_tc_site = 64;
    { EpiParams e; e.mode = EPI_SILU_ACT; e.out_act = im.a_ff.ptr; e.lda_out = kFF; e.lo_off_out = im.a_ff.lo_off;
      RUN_GEMM(im.a_ln, w.ff1_1, M, nullptr, e); }
    g_tc_site = 128;
    res = residual_gemm(im.a_ff, w.ff1_2, 0.5f);
    // self-attention
    AcacheOut ac{};
    if (im.acache) {
      ac.ring = (char*)im.acache + (size_t)l * im.ring_layer_elems * kv_elem; ac.is_f32 = split ? 1 : 0;
      ac.row_entry = b.row_entry; ac.row_pos = b.row_pos; ac.entry_slot = b.slot; ac.entry_head = b.head;
    }
    const bool fuse_att = fuse_ln && res.part == nullptr;
    if (!fuse_att) { launch_layernorm(im.x, M, w.n_att_g, w.n_att_b, nullptr, nullptr, 0, im.a_ln.out(), (im.acache && !lf) ? &ac : nullptr, st_, &res); ++launches_; }
    g_tc_site = 256;
    if (lf) attention_whole_utterance(l, w, *lf);
    else attention_streaming(l, w, kr, vr, im.acache ? &ac : nullptr, fuse_att);
    g_tc_site = 512;
    res = residual_gemm(im.a_ln, w.out, 1.0f);
    // convolution module
    const bool fuse_conv = fuse_ln && res.part == nullptr;
    if (!fuse_conv) { launch_layernorm(im.x, M, w.n_conv_g, w.n_conv_b, nullptr, nullptr, 0, im.a_ln.out(), nullptr, st_, &res); ++launches_; }
    g_tc_site = 1024;
    { EpiParams e; e.mode = EPI_GLU_F32; e.out_f32 = im.cglu; e.ldo = kDModel;
      if (!split) { e.out_act = reinterpret_cast<__nv_bfloat16*>(im.cglu); e.lda_out = kDModel; }      // bf16 mode: bf16 elements in the same buffer
      if (fuse_conv) gemm_ln(w.n_conv_g, w.n_conv_b, nullptr, w.pw1, e);
      else RUN_GEMM(im.a_ln, w.pw1, M, nullptr, e); }
    if (lf) {      // whole utterance: symmetric (4,4) zero padding, no time cache
      LfDwConvArgs a; a.c = split ? im.cglu : nullptr; a.c_bf16 = split ? nullptr : reinterpret_cast<const __nv_bfloat16*>(im.cglu);
      a.w = w.dw_w; a.bias = w.dw_b; a.out = im.a_ln.out(); a.M = M;
      launch_lf_dwconv(b, a, st_); ++launches_;
    } else {
      DwConvArgs a; a.c = split ? im.cglu : nullptr; a.c_bf16 = split ? nullptr : reinterpret_cast<const __nv_bfloat16*>(im.cglu);
      a.cache_tm = im.cache_tm + (size_t)l * kDModel * kTimeCtx; a.slot_stride = (long long)L_ * kDModel * kTimeCtx;
      a.w = w.dw_w; a.wt = w.dw_wt; a.bias = w.dw_b; a.out = im.a_ln.out();
      launch_dwconv(b, a, st_); ++launches_;
    }
    g_tc_site = 2048;
    res = residual_gemm(im.a_ln, w.pw2, 1.0f);
    // FFN 2
    const bool fuse_ff2 = fuse_ln && res.part == nullptr;
    if (!fuse_ff2) { launch_layernorm(im.x, M, w.n_ff2_g, w.n_ff2_b, nullptr, nullptr, 0, im.a_ln.out(), nullptr, st_, &res); ++launches_; }
    g_tc_site = 64;
    { EpiParams e; e.mode = EPI_SILU_ACT; e.out_act = im.a_ff.ptr; e.lda_out = kFF; e.lo_off_out = im.a_ff.lo_off;
      if (fuse_ff2) gemm_ln(w.n_ff2_g, w.n_ff2_b, nullptr, w.ff2_1, e);
      else RUN_GEMM(im.a_ln, w.ff2_1, M, nullptr, e); }
    g_tc_site = 128;
    res = residual_gemm(im.a_ff, w.ff2_2, 0.5f);
    // norm_out (+ next layer's norm_feed_forward1; after the last layer: operand of the joint's encoder projection)
    const bool last = l + 1 == L_;
    launch_layernorm(im.x, M, w.n_out_g, w.n_out_b, last ? nullptr : im.layers[l + 1].n_ff1_g, last ? nullptr : im.layers[l + 1].n_ff1_b,
                     1, last ? im.a_xf.out() : im.a_ln.out(), nullptr, st_, &res); ++launches_;
  }
  if (!lf) { launch_gather_output(b, im.x, im.enc_out, b.max_tenc, st_); ++launches_; }
}

void Engine::run_predictor_pass(const DecodeDev& d) {
  Impl& im = *im_;
  launch_pred_input(d, st_); ++launches_;
  { EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.gates; e.ldo = 4 * kPredH; e.bias = im.lstm_b[0];
    RUN_GEMM(im.a_pred, im.lstm[0], d.B, d.m_pred, e); }
  launch_lstm_cell(d, 0, st_); ++launches_;
  { EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.gates; e.ldo = 4 * kPredH; e.bias = im.lstm_b[1];
    RUN_GEMM(im.a_pred, im.lstm[1], d.B, d.m_pred, e); }
  launch_lstm_cell(d, 1, st_); ++launches_;
  { EpiParams e; e.mode = EPI_BIAS_ROWMAP_F32; e.out_f32 = im.pred_proj; e.ldo = kJointH; e.bias = im.joint_pred_b; e.row_map = d.pred_rowmap;
    RUN_GEMM(im.a_g, im.joint_pred, d.B, d.m_pred, e); }
}

static DecodeDev make_decode_dev(Engine::Impl& im, const EngineOptions& opt, int B, const int* slot, const int* row_off, const int* t_enc);

// Device-side description of one batched decode (no launches).
DecodeDev Engine::decode_setup(const BatchDev& b, const int* slots, int* steps, int max_steps, const float* enc_proj_rows) {
  Impl& im = *im_;
  DecodeDev d = make_decode_dev(im, opt_, b.B, slots ? slots : b.slot, b.row_off, im.batch_ints + 10 * im.Bcap);
  if (enc_proj_rows) d.enc_proj = enc_proj_rows;
  d.max_steps = b.max_tenc > kValidOut ? kMaxStepsOffline : kMaxStepsPerChunk;
  if (steps) { d.steps = steps; d.max_steps = max_steps; }
  d.fused_argmax = (opt_.gemm_backend == 2 || (opt_.gemm_backend == 0 && b.B > 16)) && tc_mask() < 0 ? 1 : 0;
  return d;
}

// joint encoder projection for every packed row, E = joint.enc(x) + bias (skipped when the rows were projected earlier: deferred
// decode), and the per-entry decode state
void Engine::decode_prologue(const BatchDev& b, const DecodeDev& d, bool project) {
  Impl& im = *im_;
  g_tc_site = 16;
  if (project) {
    EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.enc_proj; e.ldo = kJointH; e.bias = im.joint_enc_b;
    RUN_GEMM(im.a_xf, im.joint_enc, b.M, nullptr, e);
  }
  launch_decode_begin(d, st_); ++launches_;
}

// One symbol for every still-active entry: joint -> greedy selection -> TDT advance -> predictor step for the entries that emitted.
void Engine::decode_iteration(const BatchDev& b, const DecodeDev& d, int host_poll_slot) {
  Impl& im = *im_;
  g_tc_site = 16;
  launch_decode_iter_reset(d, st_); ++launches_;
  launch_joint_hidden(d, st_); ++launches_;
  if (d.fused_argmax) {      // tensor-core joint: greedy selection fused into the epilogue, logits never leave the SM
    EpiParams e; e.mode = EPI_ARGMAX; e.bias = im.joint_out_b; e.part_val = im.part_val; e.part_idx = im.part_idx;
    e.dur_out = im.dur_logits; e.blank_penalty = opt_.blank_penalty;
    RUN_GEMM(im.a_hid, im.joint_out, b.B, d.m_joint, e);
  } else {
    EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.logits; e.ldo = kJointOut; e.bias = im.joint_out_b;
    RUN_GEMM(im.a_hid, im.joint_out, b.B, d.m_joint, e);
  }
  launch_tdt_select(d, st_); ++launches_;
  if (host_poll_slot >= 0) {
    PKB_CUDA(cudaMemcpyAsync(im.counters_host + 2 * host_poll_slot, im.counters, sizeof(int), cudaMemcpyDeviceToHost, st_));
    PKB_CUDA(cudaEventRecord(im.dec_events[host_poll_slot], st_));
  }
  run_predictor_pass(d);
}

// The decode loop of a whole-utterance call as a one-node CUDA graph: a device-side WHILE around one decode iteration (tens of
// thousands of passes for an hour of audio: launched one by one each pass costs ~10 launches and a host poll).  Built per call
// (capture + instantiation ~1 ms), destroyed by the caller after the stream has been synchronised.  Returns false (nothing
// enqueued) when graphs are off or the build fails.
bool Engine::run_decode_loop_graph(const BatchDev& b, DecodeDev d) {
  Impl& im = *im_;
  if (im.graph_mode <= 0) return false;
  cudaGraph_t g = nullptr;
  cudaGraphExec_t exec = nullptr;
  const long long l0 = launches_;
  bool capturing = false;
  try {
    PKB_CUDA(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle handle;
    PKB_CUDA(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams cp{};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t n_loop = nullptr;
    PKB_CUDA(cudaGraphAddNode(&n_loop, g, nullptr, 0, &cp));
    d.loop_handle = (unsigned long long)handle;
    graph_pdl_suppressed() = true;
    PKB_CUDA(cudaStreamBeginCaptureToGraph(st_, cp.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    decode_iteration(b, d, -1);
    capturing = false;
    cudaGraph_t body_out = nullptr;
    PKB_CUDA(cudaStreamEndCapture(st_, &body_out));
    graph_pdl_suppressed() = false;
    im.lf_loop_body_launches = (int)(launches_ - l0);
    launches_ = l0;
    PKB_CUDA(cudaGraphInstantiate(&exec, g, 0));
    PKB_CUDA(cudaGraphLaunch(exec, st_));
    im.lf_loop_graph = g;
    im.lf_loop_exec = exec;
    return true;
  } catch (const std::exception& ex) {
    graph_pdl_suppressed() = false;
    if (capturing) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(st_, &junk); }
    if (exec) cudaGraphExecDestroy(exec);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    launches_ = l0;
    fprintf(stderr, "[parakeet_b200] decode-loop graph failed (%s); continuing launch by launch\n", ex.what());
    return false;
  }
}

// after the stream has been synchronised: kernels run by the loop graph, its algorithmic bytes (profile mode), release
void Engine::finish_decode_loop_graph(long long passes) {
  Impl& im = *im_;
  if (im.lf_persist_active) {      // the loop ran as one cooperative kernel
    im.lf_persist_active = false;
    unsigned st[7] = {0, 0, 0, 0, 0, 0, 0};
    PKB_CUDA(cudaMemcpy(st, im.pd_bar, sizeof(st), cudaMemcpyDeviceToHost));
    PKB_CHECK(st[1] == 0, "persistent decode loop: a grid barrier timed out");
    launches_ += 1;
    if (im.profile && im.lf_loop_prof_idx >= 0)
      im.prof_flops[im.lf_loop_prof_idx] =
          (double)st[6] * ((double)kJointOut * kJointH * 2.0 + kJointOut * 4.0 + (double)im.lf_loop_B * (2.0 * kJointH * 4.0 + 12.0));
    im.lf_loop_prof_idx = -1;
    return;
  }
  if (!im.lf_loop_exec) return;
  launches_ += passes * im.lf_loop_body_launches;
  if (im.profile && im.lf_loop_prof_idx >= 0)
    im.prof_flops[im.lf_loop_prof_idx] =
        (double)passes * ((double)kJointOut * kJointH * 2.0 + kJointOut * 4.0 + (double)im.lf_loop_B * (2.0 * kJointH * 4.0 + 12.0));
  cudaGraphExecDestroy(im.lf_loop_exec);
  cudaGraphDestroy(im.lf_loop_graph);
  im.lf_loop_exec = nullptr; im.lf_loop_graph = nullptr; im.lf_loop_prof_idx = -1;
}

void Engine::run_decode(const BatchDev& b, const int* slots, int* steps, int max_steps, const float* enc_proj_rows) {
  Impl& im = *im_;
  const DecodeDev d = decode_setup(b, slots, steps, max_steps, enc_proj_rows);
  decode_prologue(b, d, enc_proj_rows == nullptr);
  const int max_iters = b.max_tenc * (kMaxSymbols + 1) + 2;
  const int prof_i = prof_begin(3, 0.0);      // the whole loop; its algorithmic bytes are known when it ends
  // whole-utterance decode of a few utterances: the whole loop as ONE cooperative kernel with the decoder weights resident in shared
  // memory (PARAKEET_B200_DECODE_PERSIST=0: WHILE-graph / launch chain instead)
  static const bool persist_allowed = [] { const char* v = getenv("PARAKEET_B200_DECODE_PERSIST"); return !(v && v[0] == '0'); }();
  if (persist_allowed && steps != nullptr && b.max_tenc > 4 * kMaxTq && b.B <= kPersistMaxB && b.B <= sm_count_) {
    if (!im.pd_bar) {
      im.pd_bar = dev_alloc<unsigned>(8);
      im.pd_part_val = dev_alloc<float>((size_t)kPersistMaxB * sm_count_);
      im.pd_part_idx = dev_alloc<int>((size_t)kPersistMaxB * sm_count_);
      im.pd_dur = dev_alloc<float>((size_t)kPersistMaxB * kNDur);
      im.pd_xin = dev_alloc<float>((size_t)kPersistMaxB * 2 * kPredH);
      im.pd_x1 = dev_alloc<float>((size_t)kPersistMaxB * 2 * kPredH);
      im.pd_gvec = dev_alloc<float>((size_t)kPersistMaxB * kPredH);
      for (void* p : {(void*)im.pd_bar, (void*)im.pd_part_val, (void*)im.pd_part_idx, (void*)im.pd_dur, (void*)im.pd_xin, (void*)im.pd_x1, (void*)im.pd_gvec})
        im.dev_allocs.push_back(p);      // released with the engine
    }
    PKB_CUDA(cudaMemsetAsync(im.pd_bar, 0, 8 * sizeof(unsigned), st_));
    DecPersistArgs a{};
    a.d = d;
    a.w_out = im.joint_out.w; a.b_out = im.joint_out_b;
    a.w_l0 = im.lstm[0].w; a.b_l0 = im.lstm_b[0];
    a.w_l1 = im.lstm[1].w; a.b_l1 = im.lstm_b[1];
    a.w_jp = im.joint_pred.w; a.b_jp = im.joint_pred_b;
    a.bar = im.pd_bar; a.err = reinterpret_cast<int*>(im.pd_bar + 1); a.flags = reinterpret_cast<int*>(im.pd_bar + 2);
    a.passes_out = reinterpret_cast<int*>(im.pd_bar + 6);
    a.part_val = im.pd_part_val; a.part_idx = im.pd_part_idx; a.dur = im.pd_dur; a.xin = im.pd_xin; a.x1 = im.pd_x1; a.gvec = im.pd_gvec;
    a.max_passes = max_iters;
    if (launch_decode_persistent(a, sm_count_, st_)) {
      im.lf_persist_active = true;
      im.lf_loop_prof_idx = prof_i;
      im.lf_loop_B = b.B;
      if (prof_i >= 0) { im.prof_flops[prof_i] = 0.0; prof_end(prof_i); }
      return;
    }
  }
  if (steps != nullptr && b.max_tenc > 4 * kMaxTq && run_decode_loop_graph(b, d)) {
    // whole-utterance decode, device-side loop: passes (and with them the algorithmic bytes) are known once the traces are back
    im.lf_loop_prof_idx = prof_i;
    im.lf_loop_B = b.B;
    if (prof_i >= 0) { im.prof_flops[prof_i] = 0.0; prof_end(prof_i); }
    return;
  }
  int iters_done = 0;
  for (int it = 0; it < max_iters; ++it) {
    ++iters_done;
    // The host only needs "is anybody still active?".  Iteration it+1 is enqueued BEFORE the answer of iteration it is awaited,
    // so the GPU never idles on the round trip; the one iteration enqueued after the batch has finished sees m_joint == 0 and
    // m_pred == 0 and does nothing.
    decode_iteration(b, d, it & 1);
    if (it >= 1) {
      PKB_CUDA(cudaEventSynchronize(im.dec_events[(it - 1) & 1]));
      if (im.counters_host[2 * ((it - 1) & 1)] == 0) break;
    }
  }
  if (prof_i >= 0) {
    // per working iteration (the last one enqueued is the speculative no-op): the output layer of the joint, [8198,640] bf16 +
    // bias (its two input projections are evaluated once per chunk / per emitted token, not per iteration), plus per stream the
    // encoder-projection row, the predictor-projection row and the (token, duration) record; emitting iterations also stream
    // the LSTM weights (13.1 MB) -- the host does not see which ones emit, so those bytes are left out (lower bound)
    const double per_iter = (double)kJointOut * kJointH * 2.0 + kJointOut * 4.0 + (double)b.B * (2.0 * kJointH * 4.0 + 12.0);
    im.prof_flops[prof_i] = per_iter * std::max(iters_done - 1, 1);
    prof_end(prof_i);
  }
}

// ------------------------------------------------------------------------------------------------ whole-step CUDA graph
// A streaming step is a chain of ~400 dependent launches (15 per conformer layer + ~10 per decoded symbol), each a few microseconds of
// work at small per-GPU batches: launched one by one, the step sits on a floor of launch latencies and of one host round trip per
// decoded symbol.  Steps of one shape (same number of entries, rows, frames) launch exactly the same kernels with the same
// arguments -- everything that changes from step to step (slots, ring heads, cache lengths, frame offsets) is read from the batch
// descriptor in device memory -- so the second step of a shape is captured into a CUDA graph and every later one replays it:
//   [ child graph: encoder chunk + joint.enc projection + decode_begin ] -> [ WHILE node: one decode iteration per pass ]
// The WHILE node's condition is set on the device (lstm_cell_kernel, layer 1: "some entry is still active"), so the decode loop
// needs no host round trip at all.  PARAKEET_B200_GRAPH=0 keeps the launch-by-launch path; PARAKEET_B200_GRAPH_PDL=0 captures the
// encoder part without programmatic-dependent-launch edges (the loop body never uses them).  Any failure while building a graph
// falls back to the launch-by-launch path for good (one warning on stderr).
Engine::StepGraph* Engine::step_graph(const BatchDev& b) {
  Impl& im = *im_;
  if (im.graph_mode <= 0 || im.profile) return nullptr;
  const std::array<int, 6> key{b.B, b.M, b.sumT2, b.sumT3, b.max_Tq, b.max_tenc};
  auto it = im.graphs.find(key);
  if (it != im.graphs.end()) return it->second.get();
  if (++im.graph_seen[key] < 2) return nullptr;      // the first step of a shape runs launch by launch (it also sets the kernels' attributes)
  if (im.graphs.size() >= 32) {                      // bounded cache: drop everything when many distinct shapes have been seen
    for (auto& kv : im.graphs) {
      cudaGraphExecDestroy(kv.second->exec); cudaGraphDestroy(kv.second->graph);
      cudaEventDestroy(kv.second->ev_loop0); cudaEventDestroy(kv.second->ev_loop1);
    }
    im.graphs.clear();
  }
  std::unique_ptr<StepGraph> sg(new StepGraph());
  cudaGraph_t g_enc = nullptr;
  const long long l0 = launches_;
  bool capturing = false;
  try {
    DecodeDev d = decode_setup(b, nullptr, nullptr, 0, nullptr);
    // ---- part 1: encoder chunk + decode prologue
    graph_pdl_suppressed() = im.graph_pdl == 0;
    PKB_CUDA(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    run_encoder(b);
    decode_prologue(b, d, true);
    capturing = false;
    PKB_CUDA(cudaStreamEndCapture(st_, &g_enc));
    graph_pdl_suppressed() = false;
    sg->fixed_launches = launches_ - l0;
    // ---- the step graph: part 1 as a child graph, then the WHILE node
    PKB_CUDA(cudaGraphCreate(&sg->graph, 0));
    cudaGraphNode_t n_enc = nullptr, n_loop = nullptr, n_ev0 = nullptr, n_ev1 = nullptr;
    PKB_CUDA(cudaGraphAddChildGraphNode(&n_enc, sg->graph, nullptr, 0, g_enc));
    PKB_CUDA(cudaEventCreate(&sg->ev_loop0));
    PKB_CUDA(cudaEventCreate(&sg->ev_loop1));
    PKB_CUDA(cudaGraphAddEventRecordNode(&n_ev0, sg->graph, &n_enc, 1, sg->ev_loop0));
    cudaGraphConditionalHandle handle;
    PKB_CUDA(cudaGraphConditionalHandleCreate(&handle, sg->graph, 1, cudaGraphCondAssignDefault));      // the first pass always runs
    cudaGraphNodeParams cp{};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    PKB_CUDA(cudaGraphAddNode(&n_loop, sg->graph, &n_ev0, 1, &cp));
    PKB_CUDA(cudaGraphAddEventRecordNode(&n_ev1, sg->graph, &n_loop, 1, sg->ev_loop1));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    // ---- part 2: the loop body, captured straight into the node's body graph (plain edges)
    d.loop_handle = (unsigned long long)handle;
    const long long l1 = launches_;
    graph_pdl_suppressed() = true;
    PKB_CUDA(cudaStreamBeginCaptureToGraph(st_, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    decode_iteration(b, d, -1);
    capturing = false;
    cudaGraph_t body_out = nullptr;
    PKB_CUDA(cudaStreamEndCapture(st_, &body_out));
    graph_pdl_suppressed() = false;
    sg->body_launches = (int)(launches_ - l1);
    PKB_CUDA(cudaGraphInstantiate(&sg->exec, sg->graph, 0));
    cudaGraphDestroy(g_enc);      // the child node holds its own copy
    launches_ = l0;               // nothing was executed while capturing
  } catch (const std::exception& ex) {
    graph_pdl_suppressed() = false;
    if (capturing) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(st_, &junk); if (junk && junk != g_enc) cudaGraphDestroy(junk); }
    if (g_enc) cudaGraphDestroy(g_enc);
    if (sg->exec) cudaGraphExecDestroy(sg->exec);
    if (sg->graph) cudaGraphDestroy(sg->graph);
    if (sg->ev_loop0) cudaEventDestroy(sg->ev_loop0);
    if (sg->ev_loop1) cudaEventDestroy(sg->ev_loop1);
    cudaGetLastError();
    launches_ = l0;
    if (im.graph_pdl != 0) {
      fprintf(stderr, "[parakeet_b200] step graph with programmatic edges failed (%s); retrying with plain edges\n", ex.what());
      im.graph_pdl = 0;
      --im.graph_seen[key];
      return step_graph(b);
    }
    fprintf(stderr, "[parakeet_b200] CUDA-graph capture of the step failed (%s); continuing launch by launch\n", ex.what());
    im.graph_mode = 0;
    return nullptr;
  }
  StepGraph* out = sg.get();
  im.graphs[key] = std::move(sg);
  return out;
}

static DecodeDev make_decode_dev(Engine::Impl& im, const EngineOptions& opt, int B, const int* slot, const int* row_off, const int* t_enc) {
  DecodeDev d;
  d.B = B; d.max_symbols = kMaxSymbols; d.punct_suppress = opt.punct_suppress; d.blank_penalty = opt.blank_penalty;
  d.slot = slot; d.row_off = row_off; d.t_enc = t_enc;
  d.t_cur = im.t_cur; d.n_sym = im.n_sym; d.active = im.active; d.emit_tok = im.emit_tok; d.pred_rowmap = im.pred_rowmap;
  d.n_steps = im.n_steps; d.steps = im.steps; d.n_active = im.counters; d.m_pred = im.counters + 1; d.m_joint = im.counters + 2;
  d.enc_proj = im.enc_proj; d.pred_proj = im.pred_proj; d.logits = im.logits; d.gates = im.gates; d.embed = im.embed;
  d.part_val = im.part_val; d.part_idx = im.part_idx; d.dur_logits = im.dur_logits;
  d.punct_bits = im.punct_bits; d.pred_h = im.pred_h; d.pred_c = im.pred_c; d.pred_g = im.pred_g; d.n_emitted = im.n_emitted;
  d.y_id = im.y_id; d.act_hidden = im.a_hid.out(); d.act_pred = im.a_pred.out(); d.act_g = im.a_g.out();
  return d;
}

// reset_utterance priming (parakeet_trt.cpp:1886-1942): predictor on <|startoftranscript|> then <|en|>; blank if neither exists
void Engine::prime_streams(const std::vector<int>& sids) {
  Impl& im = *im_;
  const int B = (int)sids.size();
  if (B == 0) return;
  PKB_CHECK(B <= im.Bcap, "too many streams to prime");
  PKB_CUDA(cudaStreamSynchronize(st_));
  int* h = im.batch_ints_host;
  for (int i = 0; i < B; ++i) h[i] = streams_[sids[i]]->slot;
  PKB_CUDA(cudaMemcpyAsync(im.batch_ints, h, B * sizeof(int), cudaMemcpyHostToDevice, st_));
  DecodeDev d = make_decode_dev(im, opt_, B, im.batch_ints, nullptr, nullptr);
  std::vector<int> seq;
  if (tok_start_ >= 0) seq.push_back(tok_start_);
  if (tok_lang_ >= 0) seq.push_back(tok_lang_);
  if (seq.empty()) seq.push_back(kBlank);
  for (int tok : seq) {
    PKB_CUDA(cudaStreamSynchronize(st_));
    for (int i = 0; i < B; ++i) im.res_host[i] = tok;
    PKB_CUDA(cudaMemcpyAsync(im.force_toks, im.res_host, B * sizeof(int), cudaMemcpyHostToDevice, st_));
    launch_decode_iter_reset(d, st_); ++launches_;
    launch_force_token(d, im.force_toks, st_); ++launches_;
    run_predictor_pass(d);
  }
  // y_id = last primed token
  std::vector<int> y(B, seq.back());
  for (int i = 0; i < B; ++i)
    PKB_CUDA(cudaMemcpyAsync(im.y_id + streams_[sids[i]]->slot, &y[i], sizeof(int), cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
}

void Engine::run_batch(const std::vector<Entry>& entries, float* enc_out_host) {
  Impl& im = *im_;
  if (entries.empty()) return;
  if (enc_out_host == nullptr) {
    std::vector<int> fresh;
    for (const Entry& e : entries)
      if (streams_[e.sid]->needs_prime) { fresh.push_back(e.sid); streams_[e.sid]->needs_prime = false; }
    prime_streams(fresh);
  }
  for (auto& sp : streams_) sp->last_entry = -1;      // enc_out is about to be overwritten
  flush_feature_stage(false);                          // staged feature pushes -> feature rings (same stream: ordered before the pass)
  const BatchDev b = upload_batch(entries);
  const int max_steps = b.max_tenc > kValidOut ? kMaxStepsOffline : kMaxStepsPerChunk;
  const bool decode = enc_out_host == nullptr;
  StepGraph* sg = decode ? step_graph(b) : nullptr;
  if (sg) PKB_CUDA(cudaGraphLaunch(sg->exec, st_));
  else run_encoder(b);
  if (decode) {
    if (!sg) run_decode(b);
    // [Bcap] step counts, then [B][max_steps][3] records
    PKB_CUDA(cudaMemcpyAsync(im.res_host, im.n_steps, ((size_t)im.Bcap + (size_t)b.B * max_steps * 3) * sizeof(int),
                             cudaMemcpyDeviceToHost, st_));
  } else {
    PKB_CUDA(cudaMemcpyAsync(enc_out_host, im.enc_out, (size_t)b.B * kDModel * b.max_tenc * sizeof(float), cudaMemcpyDeviceToHost, st_));
  }
  PKB_CUDA(cudaStreamSynchronize(st_));
  im.frontend_inflight = false;
  if (im.fstage_inflight) { im.fstage_inflight = false; im.fstage_used = 0; im.fpush_n = 0; im.fpush_max_T = 0; }
  if (im.profile) profile_collect();
  const int* h = im.batch_ints_host;
  const int C = im.Bcap;
  if (sg) {      // every loop pass records one step for each entry that was still active: passes = the longest trace
    int passes = 1;
    for (int i = 0; i < b.B; ++i) passes = std::max(passes, std::min(im.res_host[i], max_steps));
    launches_ += sg->fixed_launches + (long long)passes * sg->body_launches;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sg->ev_loop0, sg->ev_loop1) == cudaSuccess) {
      im.loop_ms += ms; im.loop_passes += passes; im.loop_count += 1;
      // algorithmic bytes of a pass: see run_decode
      im.loop_bytes += (double)passes * ((double)kJointOut * kJointH * 2.0 + kJointOut * 4.0 + (double)b.B * (2.0 * kJointH * 4.0 + 12.0));
    } else {
      cudaGetLastError();
    }
  }
  for (int i = 0; i < b.B; ++i) {
    Stream& s = *streams_[entries[i].sid];
    const int Tq = h[6 * C + i];
    const int keep = Tq - kCacheDrop;
    s.last = ChunkResult();
    s.last.encoded_len = h[10 * C + i];
    s.last_entry = i; s.last_out_T = b.max_tenc;
    if (decode) {
      const int n = std::min(im.res_host[i], max_steps);
      const int* st = im.res_host + C + (size_t)i * max_steps * 3;
      for (int k = 0; k < n; ++k) {
        s.last.steps.push_back(StepRecord{st[3 * k], st[3 * k + 1], st[3 * k + 2]});
        if (st[3 * k + 1] != kBlank) { s.tokens.push_back(st[3 * k + 1]); s.token_frames.push_back((int)(s.enc_frames + st[3 * k])); }
      }
      s.enc_frames += s.last.encoded_len;
    }
    if (!s.offline) {
      s.cache_len = std::min(s.cache_len + keep, kCacheS);     // clamp(len + cache_keep_size, max=cache_len)
      s.head = (s.head + keep) % kRingCap;
    }
    s.last.cache_len_out = s.cache_len;
    s.chunks += 1;
  }
}

int Engine::step() {
  frontend_pass();
  Impl& im = *im_;
  int total = 0;
  std::vector<Entry> batch;
  int rows = 0, t3 = 0, t2 = 0;
  // A chunk leaves its stream's queue (explicit chunks) or advances its schedule (audio mode) only AFTER its batched pass has
  // succeeded: a pass that throws (bad chunk length, row capacity) leaves every chunk of the batch where it was.
  auto flush = [&]() {
    if (batch.empty()) return;
    run_batch(batch, nullptr);
    for (const Entry& e : batch) {
      Stream& s = *streams_[e.sid];
      if (s.audio_mode) s.sched_chunk += 1; else s.pending.pop_front();
    }
    total += (int)batch.size();
    batch.clear();
    rows = t3 = t2 = 0;
  };
  for (int sid = 0; sid < (int)streams_.size(); ++sid) {
    Stream& s = *streams_[sid];
    if (!s.open) continue;
    Entry e{sid, 0, 0};
    bool have = false;
    if (s.audio_mode) {
      long long lo, hi;
      sched_chunk_range(s.sched_chunk, &lo, &hi);
      if (s.frames_written >= hi) {
        e.f0 = lo;
        e.T = (int)(hi - lo);
        have = true;
      }
    } else if (!s.pending.empty()) {
      e = s.pending.front();
      have = true;
    }
    if (!have) continue;
    const int T3 = sub_len(sub_len(sub_len(e.T)));
    const int T2 = sub_len(sub_len(e.T));
    const int Tq = T3 - (s.offline ? 0 : kDropPre);
    if (rows + Tq > im.Mcap || t3 + T3 > im.T3cap || t2 + T2 > im.T2cap) flush();
    batch.push_back(e);
    rows += Tq; t3 += T3; t2 += T2;
  }
  flush();
  if (im.frontend_inflight) {      // no batched pass ran (or its last one did not follow the frontend): settle the frontend's copies
    PKB_CUDA(cudaStreamSynchronize(st_));
    im.frontend_inflight = false;
  }
  return total;
}

// ================================================================================================ tensor-level entry points
void Engine::import_state(int sid, const float* cache_ch, long long, const float* cache_tm, int cache_len) {
  Impl& im = *im_;
  PKB_CHECK(im.acache != nullptr, "state import needs contract_cache=1");
  PKB_CHECK(cache_len >= 0 && cache_len <= kCacheS, "cache_last_channel_len out of range");
  Stream& s = *streams_[sid];
  const bool split = opt_.precision == 1;
  const size_t kv_elem = split ? 4 : 2;
  s.cache_len = cache_len;
  s.head = 0;
  int hv[2] = {s.slot, 0};
  int* meta = im.meta2;
  PKB_CUDA(cudaStreamSynchronize(st_));      // meta2 / scratch reuse
  PKB_CUDA(cudaMemcpyAsync(meta, hv, sizeof(hv), cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaMemcpyAsync(im.scratch_f32, cache_ch, im.scratch_f32_elems * 4, cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaMemcpyAsync(im.cache_tm + (size_t)s.slot * L_ * kDModel * kTimeCtx, cache_tm, (size_t)L_ * kDModel * kTimeCtx * 4,
                           cudaMemcpyHostToDevice, st_));
  for (int l = 0; l < L_; ++l) {
    char* ac = (char*)im.acache + (size_t)l * im.ring_layer_elems * kv_elem;
    launch_acache_import(ac, split, meta, meta + 1, 1, im.scratch_f32 + (size_t)l * kCacheS * kDModel, 0, st_); ++launches_;
    launch_acache_to_act(ac, split, meta, meta + 1, 1, im.a_imp.out(), st_); ++launches_;
    g_tc_site = 32;
    EpiParams e; e.mode = EPI_QKV; e.n_off = kDModel; e.row_entry = im.imp_row_entry; e.row_pos = im.imp_row_pos;
    e.entry_slot = meta; e.entry_head = meta + 1;
    e.kring = (char*)im.kring + (size_t)l * im.ring_layer_elems * kv_elem;
    e.vring = (char*)im.vring + (size_t)l * im.ring_layer_elems * kv_elem;
    e.kv_f32 = split;
    e.k_natural = im.attn_mma ? 1 : 0;
    RUN_GEMM(im.a_imp, im.layers[l].kv, kCacheS, nullptr, e);
  }
  PKB_CUDA(cudaStreamSynchronize(st_));
}

void Engine::export_state(int sid, float* cache_ch, float* cache_tm) {
  Impl& im = *im_;
  PKB_CHECK(im.acache != nullptr, "state export needs contract_cache=1");
  Stream& s = *streams_[sid];
  const bool split = opt_.precision == 1;
  const size_t kv_elem = split ? 4 : 2;
  int hv[2] = {s.slot, s.head};
  int* meta = im.meta2;
  PKB_CUDA(cudaStreamSynchronize(st_));
  PKB_CUDA(cudaMemcpyAsync(meta, hv, sizeof(hv), cudaMemcpyHostToDevice, st_));
  for (int l = 0; l < L_; ++l) {
    const char* ac = (const char*)im.acache + (size_t)l * im.ring_layer_elems * kv_elem;
    launch_acache_export(ac, split, meta, meta + 1, 1, im.scratch_f32 + (size_t)l * kCacheS * kDModel, 0, st_); ++launches_;
  }
  PKB_CUDA(cudaMemcpyAsync(cache_ch, im.scratch_f32, im.scratch_f32_elems * 4, cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaMemcpyAsync(cache_tm, im.cache_tm + (size_t)s.slot * L_ * kDModel * kTimeCtx, (size_t)L_ * kDModel * kTimeCtx * 4,
                           cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
  // rows that were never filled are zeros in the contract cache (zero-initialised FIFO)
  const int invalid = kCacheS - s.cache_len;
  for (int l = 0; l < L_; ++l) memset(cache_ch + (size_t)l * kCacheS * kDModel, 0, (size_t)invalid * kDModel * 4);
}

void Engine::import_stream_state(int sid, const float* cache_ch, const float* cache_tm, int cache_len) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  import_state(sid, cache_ch, 0, cache_tm, cache_len);
}
void Engine::export_stream_state(int sid, float* cache_ch, float* cache_tm, int* cache_len) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  export_state(sid, cache_ch, cache_tm);
  *cache_len = streams_[sid]->cache_len;
}
// predictor state in the contract layout of one stream: h,c [2,640], g [640]; also refreshes the cached joint.pred(g)
void Engine::set_decoder_state(int sid, const float* h, const float* c, const float* g, int n_emitted, int y_id) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Impl& im = *im_;
  const size_t slot = streams_[sid]->slot;
  streams_[sid]->needs_prime = false;
  PKB_CUDA(cudaStreamSynchronize(st_));
  PKB_CUDA(cudaMemcpy(im.pred_h + slot * kPredL * kPredH, h, kPredL * kPredH * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(im.pred_c + slot * kPredL * kPredH, c, kPredL * kPredH * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(im.pred_g + slot * kPredH, g, kPredH * 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(im.n_emitted + slot, &n_emitted, 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaMemcpy(im.y_id + slot, &y_id, 4, cudaMemcpyHostToDevice));
  PKB_CUDA(cudaDeviceSynchronize());   // pageable H2D: staged != landed (see dev_upload)
  f32_to_act_kernel<<<(kPredH + 255) / 256, 256, 0, st_>>>(im.pred_g + slot * kPredH, kPredH, 1, 1, kPredH, im.a_g.out());
  int sl = (int)slot;
  PKB_CUDA(cudaMemcpyAsync(im.pred_rowmap, &sl, 4, cudaMemcpyHostToDevice, st_));
  g_tc_site = 16;
  { EpiParams e; e.mode = EPI_BIAS_ROWMAP_F32; e.out_f32 = im.pred_proj; e.ldo = kJointH; e.bias = im.joint_pred_b; e.row_map = im.pred_rowmap;
    RUN_GEMM(im.a_g, im.joint_pred, 1, nullptr, e); }
  PKB_CUDA(cudaStreamSynchronize(st_));
  launches_ += 1;
}
void Engine::get_decoder_state(int sid, float* h, float* c, float* g) {
  PKB_CHECK(sid >= 0 && sid < (int)streams_.size() && streams_[sid]->open, "bad stream id");
  Impl& im = *im_;
  const size_t slot = streams_[sid]->slot;
  PKB_CUDA(cudaStreamSynchronize(st_));
  PKB_CUDA(cudaMemcpy(h, im.pred_h + slot * kPredL * kPredH, kPredL * kPredH * 4, cudaMemcpyDeviceToHost));
  PKB_CUDA(cudaMemcpy(c, im.pred_c + slot * kPredL * kPredH, kPredL * kPredH * 4, cudaMemcpyDeviceToHost));
  PKB_CUDA(cudaMemcpy(g, im.pred_g + slot * kPredH, kPredH * 4, cudaMemcpyDeviceToHost));
}

// Streams borrowed by a tensor-level call: closed again on every exit path (an exception must not leak slots).
struct Engine::TempStreams {
  Engine* e;
  std::vector<int> sids;
  explicit TempStreams(Engine* eng) : e(eng) {}
  int open() { sids.push_back(e->open_stream()); return sids.back(); }
  ~TempStreams() { for (int s : sids) e->close_stream(s); }
};

void Engine::encoder_streaming_step(int B, int T, const float* audio_signal, const int64_t* length, const float* cache_last_channel,
                                    const float* cache_last_time, const int64_t* cache_last_channel_len, float* encoder_output,
                                    int64_t* encoded_lengths, float* cache_last_channel_out, float* cache_last_time_out,
                                    int64_t* cache_last_channel_len_out) {
  PKB_CHECK(B >= 1 && B <= opt_.max_streams, "encoder_streaming_step: B exceeds max_streams");
  // everything that can be rejected is rejected before a stream slot is borrowed
  const int Tq = sub_len(sub_len(sub_len(T))) - kDropPre;
  PKB_CHECK(T >= 1 && T <= 256 && Tq >= kCacheDrop && Tq <= kMaxTq, "chunk must hold 33..256 feature frames (got " + std::to_string(T) + ")");
  PKB_CHECK(im_->acache != nullptr, "encoder_streaming_step needs contract_cache=1");
  for (int i = 0; i < B; ++i) {
    PKB_CHECK(length[i] == T, "encoder_streaming_step: length must equal T for every stream");
    PKB_CHECK(cache_last_channel_len[i] >= 0 && cache_last_channel_len[i] <= kCacheS, "cache_last_channel_len out of range");
  }
  const size_t ch_stride = (size_t)L_ * kCacheS * kDModel, tm_stride = (size_t)L_ * kDModel * kTimeCtx;
  TempStreams tmp(this);
  std::vector<Entry> entries;
  for (int i = 0; i < B; ++i) {
    const int sid = tmp.open();
    import_state(sid, cache_last_channel + i * ch_stride, 0, cache_last_time + i * tm_stride, (int)cache_last_channel_len[i]);
    queue_features(sid, audio_signal + (size_t)i * kNMels * T, T);
    entries.push_back(streams_[sid]->pending.front());
    streams_[sid]->pending.pop_front();
  }
  run_batch(entries, encoder_output);
  for (int i = 0; i < B; ++i) {
    Stream& s = *streams_[tmp.sids[i]];
    encoded_lengths[i] = s.last.encoded_len;
    cache_last_channel_len_out[i] = s.cache_len;
    export_state(tmp.sids[i], cache_last_channel_out + i * ch_stride, cache_last_time_out + i * tm_stride);
  }
}

// Offline encoder at the contract layout (contract.json:67-96): audio_signal [B,128,T], length [B] (== T) ->
// encoder_output [B,1024,T_enc], encoded_lengths [B];  T <= 256 (the reference engine's profile, contract.json:284-287).
void Engine::encoder_offline_step(int B, int T, const float* audio_signal, const int64_t* length, float* encoder_output,
                                  int64_t* encoded_lengths) {
  PKB_CHECK(B >= 1 && B <= opt_.max_streams, "encoder_offline_step: B exceeds max_streams");
  PKB_CHECK(T >= 1 && T <= 256, "offline push must hold 1..256 feature frames (got " + std::to_string(T) + ")");
  for (int i = 0; i < B; ++i) PKB_CHECK(length[i] == T, "encoder_offline_step: length must equal T for every stream");
  TempStreams tmp(this);
  std::vector<Entry> entries;
  for (int i = 0; i < B; ++i) {
    const int sid = tmp.open();
    set_stream_offline(sid, true);
    queue_features(sid, audio_signal + (size_t)i * kNMels * T, T);
    entries.push_back(streams_[sid]->pending.front());
    streams_[sid]->pending.pop_front();
  }
  run_batch(entries, encoder_output);
  for (int i = 0; i < B; ++i) encoded_lengths[i] = streams_[tmp.sids[i]]->last.encoded_len;
}

// ------------------------------------------------------------------------------------------------ whole-utterance offline path
// Buffers of the long-form path are created on first use (a streaming-only server never pays for them) and owned by the Engine.
void Engine::lf_prepare(size_t total_frames, size_t steps_ints) {
  Impl& im = *im_;
  const bool split = opt_.precision == 1;
  auto own = [&](void* p) { im.dev_allocs.push_back(p); return p; };
  auto disown_free = [&](void* p) {
    if (!p) return;
    auto it = std::find(im.dev_allocs.begin(), im.dev_allocs.end(), p);
    if (it != im.dev_allocs.end()) im.dev_allocs.erase(it);
    cudaFree(p);
  };
  if (im.lf_wpos.empty()) {
    WeightsFile wf(opt_.model_dir + "/weights.bin");
    for (int l = 0; l < L_; ++l) {
      GemmW w = upload_gemm_w(wf.bf16("encoder.layers." + std::to_string(l) + ".self_attn.linear_pos.weight"), kDModel, kDModel);
      own(w.w);
      im.lf_wpos.push_back(w);
    }
    im.lf_pos_act = make_act(2 * im.Mcap, kDModel, true, st_);
    own(im.lf_pos_act.ptr);
    const size_t rows = ((size_t)im.Mcap + 127) / 128 * 128;
    im.lf_qkv = own(dev_alloc_bytes(rows * 3 * kDModel * (split ? 4 : 2)));
    im.lf_ppos = own(dev_alloc_bytes((size_t)2 * im.Mcap * kDModel * (split ? 4 : 2)));
    { const char* v = getenv("PARAKEET_B200_LF_ATTN"); if (v) im.lf_attn_kind = atoi(v); }
    if (!split && im.lf_attn_kind == 2) {
      // zero-filled once: rows / columns past the utterances are read by whole-tile TMA boxes and must hold finite values
      im.lf_q_rows = (int)rows;
      im.lf_qplanes = (__nv_bfloat16*)own(dev_alloc_bytes((size_t)2 * rows * kDModel * 2));
      PKB_CUDA(cudaMemsetAsync(im.lf_qplanes, 0, (size_t)2 * rows * kDModel * 2, st_));
      im.lf_ldv = ((long long)im.Mcap + 64LL * im.Bcap + 7) & ~7LL;
      im.lf_vt = (__nv_bfloat16*)own(dev_alloc_bytes((size_t)kDModel * im.lf_ldv * 2));
      PKB_CUDA(cudaMemsetAsync(im.lf_vt, 0, (size_t)kDModel * im.lf_ldv * 2, st_));
    }
  }
  if (total_frames > im.lf_feat_frames) {
    disown_free(im.lf_feat);
    im.lf_feat = (float*)own(dev_alloc_bytes(total_frames * kNMels * sizeof(float)));
    im.lf_feat_frames = total_frames;
  }
  if (steps_ints > im.lf_steps_ints) {
    disown_free(im.lf_steps);
    im.lf_steps = (int*)own(dev_alloc_bytes(steps_ints * sizeof(int)));
    im.lf_steps_ints = steps_ints;
  }
}

void Engine::offline_utterances(int n, const int* sids, const float* const* pcm, const size_t* n_samples, int per_feature_norm,
                                const float* const* feats, const int* T_in, int bins_major, float* const* enc_out, int decode) {
  Impl& im = *im_;
  PKB_CHECK(n >= 1 && n <= im.Bcap, "offline_utterances: 1..max_streams utterances per call");
  PKB_CHECK((pcm != nullptr) != (feats != nullptr), "offline_utterances: give either audio or features");
  PKB_CHECK(pcm ? n_samples != nullptr : T_in != nullptr, "offline_utterances: missing lengths");
  std::vector<int> T(n), f0(n);
  size_t total_frames = 0;
  for (int i = 0; i < n; ++i) {
    PKB_CHECK(sids[i] >= 0 && sids[i] < (int)streams_.size() && streams_[sids[i]]->open, "offline_utterances: bad stream id");
    const Stream& s = *streams_[sids[i]];
    PKB_CHECK(s.frames_written == 0 && s.chunks == 0 && s.pending.empty(), "offline_utterances: only on a freshly opened / reset stream");
    for (int j = 0; j < i; ++j) PKB_CHECK(sids[j] != sids[i], "offline_utterances: a stream may carry only one utterance per call");
    long long t = pcm ? (n_samples[i] >= 400 ? (long long)((n_samples[i] - 400) / 160 + 1) : 0) : T_in[i];
    PKB_CHECK(t >= 1 && t < (1 << 28), "offline_utterances: utterance needs at least one feature frame");
    T[i] = (int)t;
    f0[i] = (int)total_frames;
    total_frames += (size_t)t;
  }
  PKB_CHECK(total_frames < (1u << 30), "offline_utterances: too many frames in one call");
  // ---- batch descriptor (same fields as a streaming step; every entry is one utterance, nothing dropped, no cache)
  int* h = im.batch_ints_host;
  const int C = im.Bcap;
  PKB_CUDA(cudaStreamSynchronize(st_));
  int *slot = h, *Tf = h + C, *f0h = h + 2 * C, *T1 = h + 3 * C, *T2 = h + 4 * C, *T3 = h + 5 * C, *Tq = h + 6 * C, *qlen = h + 7 * C,
      *len = h + 8 * C, *head = h + 9 * C, *tenc = h + 10 * C, *drop = h + 11 * C, *offl = h + 12 * C;
  int* off2 = h + kNumBatchFields * C;
  int* off3 = off2 + (C + 1);
  int* roff = off3 + (C + 1);
  BatchDev b;
  b.B = n;
  off2[0] = off3[0] = roff[0] = 0;
  for (int i = 0; i < n; ++i) {
    slot[i] = 0;                              // feature base: all utterances live in the one lf_feat buffer, at frame f0
    head[i] = streams_[sids[i]]->slot;        // (the ring head is unused here: this field carries the decoder-state slot)
    Tf[i] = T[i]; f0h[i] = f0[i];
    T1[i] = sub_len(T[i]); T2[i] = sub_len(T1[i]); T3[i] = sub_len(T2[i]);
    Tq[i] = T3[i]; qlen[i] = T3[i]; len[i] = 0; tenc[i] = T3[i]; drop[i] = 0; offl[i] = 1;
    off2[i + 1] = off2[i] + T2[i]; off3[i + 1] = off3[i] + T3[i]; roff[i + 1] = roff[i] + Tq[i];
    b.max_Tq = std::max(b.max_Tq, Tq[i]);
  }
  b.max_tenc = b.max_Tq;
  b.M = roff[n]; b.sumT2 = off2[n]; b.sumT3 = off3[n];
  PKB_CHECK(b.M <= im.Mcap && b.sumT3 <= im.T3cap && b.sumT2 <= im.T2cap,
            "offline_utterances: " + std::to_string(b.M) + " encoder frames exceed the engine's row capacity (max_rows = " +
                std::to_string(im.Mcap) + ")");
  const int steps_per = b.max_tenc * (kMaxSymbols + 1);
  lf_prepare(total_frames, decode ? (size_t)n * steps_per * 3 : 0);

  // ---- features
  if (pcm) {
    size_t total_samples = 0;
    std::vector<size_t> aoff(n);
    for (int i = 0; i < n; ++i) { aoff[i] = total_samples; total_samples += (n_samples[i] + 1) & ~(size_t)1; }   // even offsets (float2 loads)
    DevTmp tmp;
    float* d_audio = tmp.alloc<float>(total_samples + 2);
    FrontSegment* d_seg = tmp.alloc<FrontSegment>(n);
    int* d_ints = tmp.alloc<int>(2 * n + 1);
    float* d_stats = tmp.alloc<float>((size_t)n * 2 * kNMels);
    std::vector<FrontSegment> segs(n);
    std::vector<int> ints(2 * n + 1);
    for (int i = 0; i < n; ++i) {
      PKB_CUDA(cudaMemcpyAsync(d_audio + aoff[i], pcm[i], n_samples[i] * sizeof(float), cudaMemcpyHostToDevice, st_));
      segs[i] = FrontSegment{(long long)aoff[i], (long long)f0[i] * kNMels, kNMels, 0, 0, -1, 0};
      ints[i] = f0[i];
      ints[n + 1 + i] = T[i];
    }
    ints[n] = (int)total_frames;
    PKB_CUDA(cudaMemcpyAsync(d_seg, segs.data(), n * sizeof(FrontSegment), cudaMemcpyHostToDevice, st_));
    PKB_CUDA(cudaMemcpyAsync(d_ints, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice, st_));
    const int pi = prof_begin(2, (double)total_frames * (kNMels * 4.0 + 160 * 4.0));
    im.frontend.logmel(d_audio, d_seg, d_ints, n, (int)total_frames, im.lf_feat, nullptr, sm_count_, st_); ++launches_;
    prof_end(pi);
    if (per_feature_norm) {
      int maxT = 0;
      for (int i = 0; i < n; ++i) maxT = std::max(maxT, T[i]);
      im.frontend.per_feature_stats(im.lf_feat, d_seg, d_ints + n + 1, n, d_stats, st_); ++launches_;
      im.frontend.apply_norm(im.lf_feat, d_seg, d_ints + n + 1, n, maxT, d_stats, st_); ++launches_;
    }
    PKB_CUDA(cudaStreamSynchronize(st_));      // the pageable sources and the host vectors above are done with
  } else {
    for (int i = 0; i < n; ++i) {
      if (bins_major) {
        DevTmp tmp;
        float* d_tmp = tmp.alloc<float>((size_t)T[i] * kNMels);
        PKB_CUDA(cudaMemcpyAsync(d_tmp, feats[i], (size_t)T[i] * kNMels * sizeof(float), cudaMemcpyHostToDevice, st_));
        im.frontend.bins_to_frames(d_tmp, T[i], im.lf_feat, (int)im.lf_feat_frames, f0[i], st_); ++launches_;
        PKB_CUDA(cudaStreamSynchronize(st_));
      } else {
        PKB_CUDA(cudaMemcpyAsync(im.lf_feat + (size_t)f0[i] * kNMels, feats[i], (size_t)T[i] * kNMels * sizeof(float), cudaMemcpyHostToDevice, st_));
      }
    }
    PKB_CUDA(cudaStreamSynchronize(st_));
  }

  // ---- predictor priming for the utterances that will be decoded (prime_streams reuses batch_ints: do it before the upload)
  if (decode) {
    std::vector<int> fresh;
    for (int i = 0; i < n; ++i)
      if (streams_[sids[i]]->needs_prime) { fresh.push_back(sids[i]); streams_[sids[i]]->needs_prime = false; }
    prime_streams(fresh);
    // prime_streams wrote slots into batch_ints_host[0..]: restore the feature-base field
    for (int i = 0; i < n; ++i) slot[i] = 0;
  }
  const size_t nints = (size_t)kNumBatchFields * C + 3 * (C + 1);
  PKB_CUDA(cudaMemcpyAsync(im.batch_ints, h, nints * sizeof(int), cudaMemcpyHostToDevice, st_));
  int* d = im.batch_ints;
  b.slot = d; b.T = d + C; b.f0 = d + 2 * C; b.T1 = d + 3 * C; b.T2 = d + 4 * C; b.T3 = d + 5 * C; b.Tq = d + 6 * C;
  b.qlen = d + 7 * C; b.len = d + 8 * C; b.head = d + 9 * C; b.drop = d + 11 * C; b.offline = d + 12 * C;
  b.off2 = d + kNumBatchFields * C; b.off3 = b.off2 + (C + 1); b.row_off = b.off3 + (C + 1);
  b.row_entry = im.row_entry; b.row_pos = im.row_pos; b.rowmap3 = im.rowmap3;

  // ---- encoder over all frames
  LongForm lf{b.max_Tq};
  launch_lf_posemb(im.lf_pos_act.out(), lf.Tm, st_); ++launches_;
  run_encoder(b, &lf);
  std::vector<float> rows;
  if (enc_out) {
    rows.resize((size_t)b.M * kDModel);
    PKB_CUDA(cudaMemcpyAsync(rows.data(), im.x, rows.size() * sizeof(float), cudaMemcpyDeviceToHost, st_));
  }
  if (decode == 2) {
    // deferred decode: only project the rows for the joint now and park them; offline_decode_pending() decodes every parked
    // utterance of the engine in ONE batched loop (a decode iteration costs the same launch chain for 4 or for 32 utterances)
    if (im.lf_store_used + (size_t)b.M > im.lf_store_cap) {
      const size_t cap = std::max(im.lf_store_cap * 2, im.lf_store_used + (size_t)b.M);
      float* p = (float*)dev_alloc_bytes(cap * kJointH * sizeof(float));
      im.dev_allocs.push_back(p);
      if (im.lf_store_used) PKB_CUDA(cudaMemcpyAsync(p, im.lf_store, im.lf_store_used * kJointH * sizeof(float), cudaMemcpyDeviceToDevice, st_));
      PKB_CUDA(cudaStreamSynchronize(st_));
      if (im.lf_store) {
        im.dev_allocs.erase(std::find(im.dev_allocs.begin(), im.dev_allocs.end(), (void*)im.lf_store));
        cudaFree(im.lf_store);
      }
      im.lf_store = p;
      im.lf_store_cap = cap;
    }
    g_tc_site = 16;
    { EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.lf_store + im.lf_store_used * kJointH; e.ldo = kJointH; e.bias = im.joint_enc_b;
      RUN_GEMM(im.a_xf, im.joint_enc, b.M, nullptr, e); }
    for (int i = 0; i < n; ++i) {
      Stream& s = *streams_[sids[i]];
      s.lf_store_off = (long long)im.lf_store_used + roff[i];
      s.lf_store_T = h[6 * C + i];
    }
    im.lf_store_used += (size_t)b.M;
  }
  // ---- greedy TDT over every frame of every utterance
  std::vector<int> counts, recs;
  if (decode == 1) {
    run_decode(b, b.head, im.lf_steps, steps_per);
    counts.resize(n);
    recs.resize((size_t)n * steps_per * 3);
    PKB_CUDA(cudaMemcpyAsync(counts.data(), im.n_steps, n * sizeof(int), cudaMemcpyDeviceToHost, st_));
    PKB_CUDA(cudaMemcpyAsync(recs.data(), im.lf_steps, recs.size() * sizeof(int), cudaMemcpyDeviceToHost, st_));
  }
  PKB_CUDA(cudaStreamSynchronize(st_));
  if (decode == 1) { long long passes = 1; for (int c : counts) passes = std::max<long long>(passes, std::min(c, steps_per)); finish_decode_loop_graph(passes); }
  if (im.profile) profile_collect();
  for (int i = 0; i < n; ++i) {
    Stream& s = *streams_[sids[i]];
    const int Te = h[6 * C + i], r0 = roff[i];
    if (enc_out && enc_out[i]) {      // contract layout [1024, T_enc]
      float* o = enc_out[i];
      for (int t = 0; t < Te; ++t)
        for (int c = 0; c < kDModel; ++c) o[(size_t)c * Te + t] = rows[(size_t)(r0 + t) * kDModel + c];
    }
    s.last = ChunkResult();
    s.last.encoded_len = Te;
    s.frames_written += T[i];
    s.chunks += 1;
    if (decode == 1) {
      const int cnt = std::min(counts[i], steps_per);
      const int* st = recs.data() + (size_t)i * steps_per * 3;
      for (int k = 0; k < cnt; ++k) {
        s.last.steps.push_back(StepRecord{st[3 * k], st[3 * k + 1], st[3 * k + 2]});
        if (st[3 * k + 1] != kBlank) { s.tokens.push_back(st[3 * k + 1]); s.token_frames.push_back((int)(s.enc_frames + st[3 * k])); }
      }
      s.enc_frames += Te;
    }
  }
}

// Decode every utterance whose encoder rows were parked by offline_utterances(decode = 2), all in one batched TDT loop.
int Engine::offline_decode_pending() {
  Impl& im = *im_;
  std::vector<int> sids;
  for (int i = 0; i < (int)streams_.size(); ++i)
    if (streams_[i]->open && streams_[i]->lf_store_off >= 0) sids.push_back(i);
  const int n = (int)sids.size();
  if (n == 0) return 0;
  PKB_CUDA(cudaStreamSynchronize(st_));
  std::vector<int> fresh;
  for (int sid : sids)
    if (streams_[sid]->needs_prime) { fresh.push_back(sid); streams_[sid]->needs_prime = false; }
  prime_streams(fresh);
  int* h = im.batch_ints_host;
  const int C = im.Bcap;
  int* head = h + 9 * C;      // decoder-state slots (see offline_utterances)
  int* tenc = h + 10 * C;
  int* roff = h + kNumBatchFields * C + 2 * (C + 1);
  BatchDev b;
  b.B = n;
  for (int i = 0; i < n; ++i) {
    const Stream& s = *streams_[sids[i]];
    PKB_CHECK(s.lf_store_off < (1ll << 31), "deferred-decode store too large");
    head[i] = s.slot; tenc[i] = s.lf_store_T; roff[i] = (int)s.lf_store_off;      // each entry's own first row (not a prefix sum here)
    b.max_tenc = std::max(b.max_tenc, s.lf_store_T);
  }
  roff[n] = 0;
  b.max_Tq = b.max_tenc;
  const size_t nints = (size_t)kNumBatchFields * C + 3 * (C + 1);
  PKB_CUDA(cudaMemcpyAsync(im.batch_ints, h, nints * sizeof(int), cudaMemcpyHostToDevice, st_));
  b.head = im.batch_ints + 9 * C;
  b.row_off = im.batch_ints + kNumBatchFields * C + 2 * (C + 1);
  const int steps_per = b.max_tenc * (kMaxSymbols + 1);
  lf_prepare(0, (size_t)n * steps_per * 3);
  run_decode(b, b.head, im.lf_steps, steps_per, im.lf_store);
  std::vector<int> counts(n), recs((size_t)n * steps_per * 3);
  PKB_CUDA(cudaMemcpyAsync(counts.data(), im.n_steps, n * sizeof(int), cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaMemcpyAsync(recs.data(), im.lf_steps, recs.size() * sizeof(int), cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
  { long long passes = 1; for (int c : counts) passes = std::max<long long>(passes, std::min(c, steps_per)); finish_decode_loop_graph(passes); }
  if (im.profile) profile_collect();
  for (int i = 0; i < n; ++i) {
    Stream& s = *streams_[sids[i]];
    const int cnt = std::min(counts[i], steps_per);
    const int* st = recs.data() + (size_t)i * steps_per * 3;
    s.last.steps.clear();
    for (int k = 0; k < cnt; ++k) {
      s.last.steps.push_back(StepRecord{st[3 * k], st[3 * k + 1], st[3 * k + 2]});
      if (st[3 * k + 1] != kBlank) { s.tokens.push_back(st[3 * k + 1]); s.token_frames.push_back((int)(s.enc_frames + st[3 * k])); }
    }
    s.enc_frames += s.lf_store_T;
    s.lf_store_off = -1; s.lf_store_T = 0;
  }
  im.lf_store_used = 0;
  return n;
}

void Engine::predictor_step(int B, const int64_t* y, const float* h, const float* c, float* g, float* h_out, float* c_out) {
  Impl& im = *im_;
  PKB_CHECK(B >= 1 && B <= opt_.max_streams, "predictor_step: B exceeds max_streams");
  TempStreams tmp(this);
  for (int i = 0; i < B; ++i) streams_[tmp.open()]->needs_prime = false;
  const std::vector<int>& sids = tmp.sids;
  PKB_CUDA(cudaStreamSynchronize(st_));
  std::vector<float> hb(kPredL * kPredH), cb(kPredL * kPredH);
  for (int i = 0; i < B; ++i) {
    for (int l = 0; l < kPredL; ++l) {   // contract layout [2,B,640] -> slot layout [2][640]
      memcpy(&hb[l * kPredH], h + ((size_t)l * B + i) * kPredH, kPredH * 4);
      memcpy(&cb[l * kPredH], c + ((size_t)l * B + i) * kPredH, kPredH * 4);
    }
    const size_t slot = streams_[sids[i]]->slot;
    PKB_CUDA(cudaMemcpy(im.pred_h + slot * kPredL * kPredH, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    PKB_CUDA(cudaMemcpy(im.pred_c + slot * kPredL * kPredH, cb.data(), cb.size() * 4, cudaMemcpyHostToDevice));
    im.batch_ints_host[i] = (int)slot;
    im.res_host[i] = (int)y[i];
  }
  PKB_CUDA(cudaDeviceSynchronize());     // pageable H2D: staged != landed (see dev_upload)
  PKB_CUDA(cudaMemcpyAsync(im.batch_ints, im.batch_ints_host, B * sizeof(int), cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaMemcpyAsync(im.force_toks, im.res_host, B * sizeof(int), cudaMemcpyHostToDevice, st_));
  DecodeDev d = make_decode_dev(im, opt_, B, im.batch_ints, nullptr, nullptr);
  launch_decode_iter_reset(d, st_); ++launches_;
  launch_force_token(d, im.force_toks, st_); ++launches_;
  run_predictor_pass(d);
  PKB_CUDA(cudaStreamSynchronize(st_));
  for (int i = 0; i < B; ++i) {
    const size_t slot = streams_[sids[i]]->slot;
    PKB_CUDA(cudaMemcpy(hb.data(), im.pred_h + slot * kPredL * kPredH, hb.size() * 4, cudaMemcpyDeviceToHost));
    PKB_CUDA(cudaMemcpy(cb.data(), im.pred_c + slot * kPredL * kPredH, cb.size() * 4, cudaMemcpyDeviceToHost));
    for (int l = 0; l < kPredL; ++l) {
      memcpy(h_out + ((size_t)l * B + i) * kPredH, &hb[l * kPredH], kPredH * 4);
      memcpy(c_out + ((size_t)l * B + i) * kPredH, &cb[l * kPredH], kPredH * 4);
    }
    PKB_CUDA(cudaMemcpy(g + (size_t)i * kPredH, im.pred_g + slot * kPredH, kPredH * 4, cudaMemcpyDeviceToHost));  // [B,640,1]
  }
}

void Engine::joint_step(int B, int T, int U, const float* enc, const float* pred, float* out) {
  Impl& im = *im_;
  const int rows = B * T * U;
  PKB_CHECK(B >= 1 && B * T <= im.a_xf.rows_cap && B * U <= im.a_g.rows_cap && rows <= im.a_hid.rows_cap,
            "joint_step: B*T*U exceeds the decode row capacity (raise max_streams)");
  PKB_CUDA(cudaStreamSynchronize(st_));
  DevTmp tmp;
  float* d_enc = tmp.upload(enc, (size_t)B * kDModel * T);
  float* d_pred = tmp.upload(pred, (size_t)B * kPredH * U);
  float* d_P = tmp.alloc<float>((size_t)B * U * kJointH);
  // enc [B,1024,T] -> operand rows (b*T+t): row stride within b is 1 (t), col stride T
  for (int b = 0; b < B; ++b) {
    ActOut a = im.a_xf.out(); a.ptr += (size_t)b * T * a.lda;
    f32_to_act_kernel<<<(T * kDModel + 255) / 256, 256, 0, st_>>>(d_enc + (size_t)b * kDModel * T, 1, T, T, kDModel, a);
    ActOut p = im.a_g.out(); p.ptr += (size_t)b * U * p.lda;
    f32_to_act_kernel<<<(U * kPredH + 255) / 256, 256, 0, st_>>>(d_pred + (size_t)b * kPredH * U, 1, U, U, kPredH, p);
    launches_ += 2;
  }
  { EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.enc_proj; e.ldo = kJointH; e.bias = im.joint_enc_b;
    RUN_GEMM(im.a_xf, im.joint_enc, B * T, nullptr, e); }
  { EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = d_P; e.ldo = kJointH; e.bias = im.joint_pred_b;
    RUN_GEMM(im.a_g, im.joint_pred, B * U, nullptr, e); }
  joint_hidden_grid_kernel<<<rows, 128, 0, st_>>>(im.enc_proj, d_P, T, U, im.a_hid.out()); ++launches_;
  { EpiParams e; e.mode = EPI_BIAS_F32; e.out_f32 = im.logits; e.ldo = kJointOut; e.bias = im.joint_out_b;
    RUN_GEMM(im.a_hid, im.joint_out, rows, nullptr, e); }
  PKB_CUDA(cudaMemcpyAsync(out, im.logits, (size_t)rows * kJointOut * 4, cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
}

size_t Engine::logmel(const float* pcm, size_t n, float* out, int per_feature_norm) {
  Impl& im = *im_;
  if (n < 400) return 0;
  const size_t T = (n - 400) / 160 + 1;
  PKB_CHECK(T < (1u << 30), "clip too long");
  DevTmp tmp;
  float* d_audio = tmp.alloc<float>(n);
  float* d_out = tmp.alloc<float>(T * kNMels);
  float* d_stats = tmp.alloc<float>(2 * kNMels);
  PKB_CUDA(cudaMemcpyAsync(d_audio, pcm, n * 4, cudaMemcpyHostToDevice, st_));
  FrontSegment sg{0, 0, kNMels, 0, 0, -1, 0};
  int prefix[2] = {0, (int)T}, frames = (int)T;
  FrontSegment* d_seg = tmp.alloc<FrontSegment>(1);
  int* d_prefix = tmp.alloc<int>(3);
  PKB_CUDA(cudaMemcpyAsync(d_seg, &sg, sizeof(sg), cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaMemcpyAsync(d_prefix, prefix, sizeof(prefix), cudaMemcpyHostToDevice, st_));
  PKB_CUDA(cudaMemcpyAsync(d_prefix + 2, &frames, sizeof(int), cudaMemcpyHostToDevice, st_));
  const int pi = prof_begin(2, (double)T * (kNMels * 4.0 + 160 * 4.0));      // 640 B of new samples read + 512 B written per frame
  im.frontend.logmel(d_audio, d_seg, d_prefix, 1, (int)T, d_out, nullptr, sm_count_, st_); ++launches_;
  prof_end(pi);
  if (per_feature_norm) {
    im.frontend.per_feature_stats(d_out, d_seg, d_prefix + 2, 1, d_stats, st_); ++launches_;
    im.frontend.apply_norm(d_out, d_seg, d_prefix + 2, 1, (int)T, d_stats, st_); ++launches_;
  }
  PKB_CUDA(cudaMemcpyAsync(out, d_out, T * kNMels * 4, cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
  if (im.profile) profile_collect();
  return T;
}

void Engine::gemm_test(int backend, int M, int N, int K, const float* A, const uint16_t* W_bits, float* C, int epi_silu) {
  const bool split = opt_.precision == 1;
  PKB_CUDA(cudaStreamSynchronize(st_));
  ActBuf a = make_act(M, K, split, st_);
  GemmW w = upload_gemm_w(std::vector<uint16_t>(W_bits, W_bits + (size_t)N * K), N, K);
  float* d_A = dev_upload(std::vector<float>(A, A + (size_t)M * K));
  float* d_C = dev_alloc<float>((size_t)M * N);
  const long long n = (long long)M * K;
  f32_to_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st_>>>(d_A, K, 1, M, K, a.out());
  GemmArgs g;
  g.A = a.ptr; g.lda = K; g.a_lo_off = a.lo_off; g.W = w.w; g.M = M; g.N = N; g.K = K;
  g.epi.mode = EPI_F32; g.epi.out_f32 = d_C; g.epi.ldo = N;
  g.map_w32 = &w.map32;
  (void)epi_silu;
  if (backend >= 1) {     // 1: heuristic, 2: 128-wide tiles, 3: 256-wide tiles, 4: CTA-pair 256 x 256 tiles (cta_group::2)
    PKB_CHECK(gemm_tc_supported(g), "gemm_test: shape not supported by the tensor-core backend");
    gemm_tc_set_bn(backend == 2 ? 128 : backend == 3 ? 256 : backend == 4 ? 512 : 0);
    gemm_tc(g, a.map, w.map, st_);
    gemm_tc_set_bn(0);
  } else {
    gemm_simt(g, st_);
  }
  ++launches_;
  PKB_CUDA(cudaMemcpyAsync(C, d_C, (size_t)M * N * 4, cudaMemcpyDeviceToHost, st_));
  PKB_CUDA(cudaStreamSynchronize(st_));
  cudaFree(a.ptr); cudaFree(w.w); cudaFree(d_A); cudaFree(d_C);
}

}  // namespace pkb
