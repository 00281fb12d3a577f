// engine.h -- per-GPU batched streaming engine behind the C ABI.
//
// The reference drives one stream per session through three TensorRT engines with the decode loop on the host
// (/root/reference/cpp/src/parakeet_trt.cpp:1557-1665, 1967-3858).  Here one Engine owns the weights of one GPU and a table of
// stream slots; `step()` advances every stream that has a pending chunk in ONE batched pass
// (frontend -> FastConformer chunk -> TDT decode), and the legacy `ParakeetSession` is a one-stream view of it.
#pragma once
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "dec_kernels.cuh"
#include "enc_kernels.cuh"
#include "frontend.h"
#include "gemm.h"
#include "vocab.h"
#include "weights_file.h"

namespace pkb {

struct EngineOptions {
  std::string model_dir;
  int device_id = 0;
  int max_streams = 1;
  int precision = 0;        // 0 = bf16 operands (fast); 1 = split bf16 hi+lo operands, f32 K/V rings (fp32-grade)
  int gemm_backend = 0;     // 0 = auto (tcgen05 for M > 16, weight-streaming SIMT below); 1 = SIMT only; 2 = tcgen05 always
  int contract_cache = 1;   // keep the pre-projection cache_last_channel ring (needed by state export)
  int punct_suppress = 1;   // PARAKEET_DISABLE_PUNCT_SUPPRESSION inverse
  float blank_penalty = 0.0f;
  int max_rows = 0;         // packed encoder rows per batched pass (0: max(64, 8*max_streams))
};

struct StepRecord { int time_idx, token, duration; };

struct ChunkResult {          // what one chunk of one stream produced
  std::vector<StepRecord> steps;
  int cache_len_out = 0;
  int encoded_len = 0;
};

class Engine {
 public:
  explicit Engine(const EngineOptions& opt);
  ~Engine();
  Engine(const Engine&) = delete;

  int open_stream();                       // -> stream id, or throws when the table is full
  void close_stream(int sid);
  void reset_stream(int sid);              // zero caches + predictor state, prime with <|startoftranscript|>, <|en|>

  // Legacy-ABI granularity: these T frames ([128,T] bins-major, host) form exactly one encoder chunk.
  void queue_features(int sid, const float* feats_bins_major, int T);
  // Audio granularity: buffered; step() runs the GPU frontend and cuts chunks by the 41/57-frame schedule
  // (tools/verify_nemo/streaming_encoder_reference.py:522-550).
  void queue_audio(int sid, const float* pcm, size_t n);
  void set_feature_norm(int sid, const float* mean128, const float* std128);   // nullptrs: none
  // streaming-safe alternative: causal running mean / std per feature over the stream's own frames (GPU frontend, audio input)
  void set_feature_norm_running(int sid, bool on);
  // offline mode (the reference's non-streaming `encoder` engine): every push of <= 256 frames is encoded with full context and no
  // caches, all of its encoder frames are decoded; only valid on a fresh stream
  void set_stream_offline(int sid, bool offline);
  // batched push of `count` samples for n streams in one call; source rows `stride` floats apart, in host (pinned or
  // pageable) or device memory; lands in per-stream device audio buffers with one copy + one kernel
  void push_audio_batch(int n, const int* sids, const float* src, long long stride, int count, bool src_on_device,
                        bool internal = false);

  // Advance every stream that has a pending chunk by one chunk.  Returns the number of chunks processed.
  int step();
  // work left for step(): an explicit chunk, audio that still yields frames, or ring frames completing the next scheduled chunk
  // (drain with `while (has_pending(sid)) step();`)
  bool has_pending(int sid) const;
  void drop_pending(int sid);              // forget queued, unprocessed chunks of a stream (error recovery)

  const std::vector<int>& tokens(int sid) const;
  const ChunkResult& last_chunk(int sid) const;
  // Timestamps in the encoder timebase (80 ms per frame: 10 ms feature shift x 8 subsampling, docs/ARCHITECTURE_RUNTIME.md:46-47):
  // the absolute encoder frame each token was emitted on, and the number of encoder frames decoded so far.
  // encoder_output [1024, T] (contract layout, T = min(encoded_len, cap_T)) of the stream's chunk in the LAST batched pass; returns encoded_len
  int last_encoder_output(int sid, float* out, int cap_T);
  // run the deferred predictor priming of a freshly reset stream now; returns the token the predictor was last stepped on
  int prime_now(int sid);
  const std::vector<int>& token_frames(int sid) const;
  long long encoder_frames_done(int sid) const;
  // "Stable prefix + revision window" (MAGNOLIA_INTEGRATION_HANDOFF.md:105-135): number of leading tokens older than
  // revision_window_ms behind the live edge -- the part of the transcript a UI may commit.  Greedy TDT never rewrites an emitted
  // token, so the policy reduces to its time threshold.
  int stable_prefix(int sid, int revision_window_ms) const;
  int cache_len(int sid) const;
  long long chunks_done(int sid) const;

  // ---- per-stream state across the ABI (checkpoint / migration / functional-mode parity), contract layouts, host pointers
  // cache_last_channel [L,256,1024], cache_last_time [L,1024,4] of ONE stream
  void import_stream_state(int sid, const float* cache_ch, const float* cache_tm, int cache_len);
  void export_stream_state(int sid, float* cache_ch, float* cache_tm, int* cache_len);
  void set_decoder_state(int sid, const float* h, const float* c, const float* g, int n_emitted, int y_id);
  void get_decoder_state(int sid, float* h, float* c, float* g);

  // ---- tensor-level entry points at the contract layouts (host pointers) ----
  void encoder_streaming_step(int B, int T, const float* audio_signal, const int64_t* length, const float* cache_last_channel,
                              const float* cache_last_time, const int64_t* cache_last_channel_len, float* encoder_output,
                              int64_t* encoded_lengths, float* cache_last_channel_out, float* cache_last_time_out,
                              int64_t* cache_last_channel_len_out);
  void encoder_offline_step(int B, int T, const float* audio_signal, const int64_t* length, float* encoder_output /*[B,1024,T_enc]*/,
                            int64_t* encoded_lengths);
  // ---- whole-utterance offline path (BASELINE configs 1 and 5; the reference's `encoder` graph run with a dynamic time axis,
  // contract.json:67-96, + the TDT loop over all of its frames) ----
  // n utterances, utterance i bound to the freshly opened / reset stream sids[i].  Input per utterance: either 16 kHz PCM
  // (pcm[i], n_samples[i]; log-mel and optional whole-utterance per-feature normalisation run on the GPU) or features
  // (feats[i]: T[i] frames, bins-major [128,T] when bins_major else frames-major [T,128]).  The encoder attends over ALL frames
  // of an utterance (no caches, no 256-frame limit); sum of encoder frames <= max_rows.  enc_out[i] (optional, host) receives
  // encoder_output [1024, T_enc_i]; decode == 1 runs greedy TDT over every frame now: tokens(sid) / last_chunk(sid); decode == 2
  // defers it to offline_decode_pending().
  void offline_utterances(int n, const int* sids, const float* const* pcm, const size_t* n_samples, int per_feature_norm,
                          const float* const* feats, const int* T, int bins_major, float* const* enc_out, int decode);
  // decode == 2 in offline_utterances parks the utterances' rows; this decodes all parked utterances in one batched loop; returns how many
  int offline_decode_pending();
  void predictor_step(int B, const int64_t* y, const float* h, const float* c, float* g, float* h_out, float* c_out);
  void joint_step(int B, int T, int U, const float* enc, const float* pred, float* out);
  // GPU frontend on host buffers: pcm[n] -> frames-major [T,128]; per_feature_norm applies utterance mean/std
  size_t logmel(const float* pcm, size_t n, float* out_frames_major, int per_feature_norm);
  // standalone GEMM for validation of the tensor-core backend: C = A(f32, split or rounded per precision) * W^T
  void gemm_test(int backend, int M, int N, int K, const float* A, const uint16_t* W_bf16, float* C, int epi_silu);

  // NaN / Inf census of a stream's tensors after its last pass (the reference's nan_guard_device, parakeet_trt.cpp:913-1013):
  // stage 0 encoder_output, 1 cache_last_channel (newest rows, layer 0), 2 cache_last_time (layer 0)
  struct GuardResult { int nan_count = 0, inf_count = 0, first_nan = -1, first_inf = -1; size_t sample_n = 0, count = 0; };
  GuardResult nan_guard(int sid, int stage);

  const EngineOptions& options() const { return opt_; }
  int n_layers() const { return L_; }
  const Vocab& vocab() const { return vocab_; }
  std::string detokenize(const std::vector<int>& ids) const;
  long long kernel_launches() const { return launches_; }
  int sm_count() const { return sm_count_; }
  cudaStream_t stream() const { return st_; }
  void synchronize();
  // CUDA events on the engine's own stream (bench timing) and per-launch timing of the tcgen05 GEMM (roofline)
  int event_record();
  double event_elapsed_ms(int a, int b);
  void profile_enable(bool on);
  void profile_collect();
  void profile_read(int cls, double* ms, double* work, long long* launches);   // cls 0: tcgen05 GEMM (work = FLOPs), 1: attention, 2: frontend, 3: decode loop (bytes), 4: whole-utterance attention (FLOPs)
  void decode_loop_stats(double* ms, double* bytes, long long* passes, long long* loops, int reset);
  void set_blank_penalty(float p);         // PARAKEET_BLANK_PENALTY (parakeet_trt.cpp:3175-3178) at run time
  int graphs_built() const;                // step shapes currently held as CUDA graphs (0: launch-by-launch path)
  int prof_begin(int cls, double work);
  void prof_end(int idx);

  struct Stream;
  struct Impl;
  struct TempStreams;
  struct StepGraph;
  struct Entry { int sid; long long f0; int T; };      // f0: ABSOLUTE index of the chunk's first frame (ring index = f0 % ring capacity)
  Impl* impl();

 private:
  void compact_audio();
  void load_weights();
  void alloc_state();
  void run_batch(const std::vector<Entry>& entries, float* enc_out_host /*optional [B,1024,3]*/);
  struct LongForm { int Tm; };       // whole-utterance pass: Tm = longest utterance (centre of the relative-position table)
  void run_encoder(const BatchDev& b, const LongForm* lf = nullptr);
  // slots / steps / max_steps override the per-chunk defaults for the whole-utterance path
  void run_decode(const BatchDev& b, const int* slots = nullptr, int* steps = nullptr, int max_steps = 0, const float* enc_proj_rows = nullptr);
  DecodeDev decode_setup(const BatchDev& b, const int* slots, int* steps, int max_steps, const float* enc_proj_rows);
  void decode_prologue(const BatchDev& b, const DecodeDev& d, bool project);
  void decode_iteration(const BatchDev& b, const DecodeDev& d, int host_poll_slot);
  StepGraph* step_graph(const BatchDev& b);      // nullptr: run this step launch by launch
  bool run_decode_loop_graph(const BatchDev& b, DecodeDev d);
  void finish_decode_loop_graph(long long passes);
  void lf_prepare(size_t total_frames, size_t steps_ints);
  void run_predictor_pass(const DecodeDev& d);
  void frontend_pass();
  void flush_feature_stage(bool sync);
  void prime_streams(const std::vector<int>& sids);
  void import_state(int sid, const float* cache_ch, long long ch_stride_unused, const float* cache_tm, int cache_len);
  void export_state(int sid, float* cache_ch, float* cache_tm);
  BatchDev upload_batch(const std::vector<Entry>& entries);

  EngineOptions opt_;
  int L_ = 0;
  int sm_count_ = 148;
  cudaStream_t st_ = nullptr;
  std::unique_ptr<Impl> im_;
  std::vector<std::unique_ptr<Stream>> streams_;
  Vocab vocab_;
  std::vector<uint32_t> punct_bits_;
  int tok_start_ = -1, tok_lang_ = -1;
  long long launches_ = 0;
};

}  // namespace pkb
